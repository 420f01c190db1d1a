"""WCSPH (reference: SPH/fluid_solvers/WCSPH.py): Tait equation of state + explicit step."""
from .._native import T
from .base_solver import BaseSolver


class WCSPHSolver(BaseSolver):
    #: one entry per call of the reference's `_step` (WCSPH.py:27-45), in its order; "container." / "rigid_solver."
    #: prefixes name the collaborator the call goes to
    STEP_SEQUENCE = (
        "container.prepare_neighborhood_search", "compute_density", "compute_non_pressure_acceleration",
        "update_fluid_velocity",
        "compute_pressure", "compute_pressure_acceleration", "update_fluid_velocity", "update_fluid_position",
        "rigid_solver.step", "container.insert_object", "rigid_solver.insert_rigid_object", "renew_rigid_particle_state",
        "enforce_fluid_domain",
    )

    def __init__(self, container):
        super().__init__(container)
        self.gamma, self.stiffness = 7.0, 50000.0     # hard-coded upstream (WCSPH.py:12-13); the JSON keys are ignored

    def compute_pressure(self):
        self._run(T.WCSPH_COMPUTE_PRESSURE)

    def enforce_fluid_domain(self):
        self.enforce_domain_boundary_3D(self.container.material_fluid)

    def _step(self):
        for name in self.STEP_SEQUENCE:
            target = self
            *owners, method = name.split(".")
            for owner in owners:
                target = getattr(target, owner)
            getattr(target, method)()

    _library_step_impl = _step
