from .bullet_solver import PyBulletSolver

__all__ = ["PyBulletSolver"]
