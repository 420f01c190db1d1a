"""WCSPH (reference: SPH/fluid_solvers/WCSPH.py): Tait equation of state + explicit step."""
from .._native import T
from .base_solver import BaseSolver


class WCSPHSolver(BaseSolver):
    def __init__(self, container):
        super().__init__(container)
        self.gamma = 7.0          # hard-coded upstream (WCSPH.py:12-13); the JSON keys are ignored
        self.stiffness = 50000.0

    def compute_pressure(self):
        self._run(T.WCSPH_COMPUTE_PRESSURE)

    def _step(self):
        self.container.prepare_neighborhood_search()
        self.compute_density()
        self.compute_non_pressure_acceleration()
        self.update_fluid_velocity()

        self.compute_pressure()
        self.compute_pressure_acceleration()
        self.update_fluid_velocity()
        self.update_fluid_position()

        self.rigid_solver.step()
        self.container.insert_object()
        self.rigid_solver.insert_rigid_object()
        self.renew_rigid_particle_state()

        self.enforce_domain_boundary_3D(self.container.material_fluid)

    _library_step_impl = _step
