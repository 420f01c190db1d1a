"""Import stub: the reference imports pybullet at module level; scenes without RigidBodies never call it
(SPH/rigid_solver/bullet_solver.py:40-42,144-146)."""
DIRECT = 0


def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError("pybullet is not available offline")
    return _missing
