"""PCISPH (reference: SPH/fluid_solvers/PCISPH.py): predictive-corrective pressure loop."""
from .._native import T
from .base_solver import BaseSolver


class PCISPHSolver(BaseSolver):
    def __init__(self, container):
        super().__init__(container)
        self.max_iterations = 1000
        self.eta = 0.001
        self.last_iterations = None

    def compute_predicted_velocity(self):
        self._run(T.PCISPH_COMPUTE_PREDICTED_VELOCITY)

    def compute_predicted_position(self):
        self._run(T.PCISPH_COMPUTE_PREDICTED_POSITION)

    def compute_density_star(self):
        self._run(T.PCISPH_COMPUTE_DENSITY_STAR)

    def update_pressure(self):
        self._run(T.PCISPH_UPDATE_PRESSURE)

    def compute_temp_pressure_acceleration(self):
        self._run(T.PCISPH_COMPUTE_TEMP_PRESSURE_ACCELERATION)

    def refine(self):
        num_itr = 0
        while num_itr < self.max_iterations:
            self.compute_density_star()
            self.update_pressure()
            self.compute_temp_pressure_acceleration()
            self.compute_predicted_velocity()
            self.compute_predicted_position()
            num_itr += 1
            if self.container.density_error[None] < self.eta:
                break
        self.last_iterations = (num_itr, self.container.density_error[None])
        return num_itr

    def compute_pcisph_k(self):
        self._run(T.PCISPH_COMPUTE_K)

    def init_step(self):
        self._run(T.PCISPH_INIT_STEP)

    def _step(self):
        self.container.prepare_neighborhood_search()
        self.compute_density()
        self.compute_non_pressure_acceleration()
        self.init_step()
        self.refine()

        self.update_fluid_velocity()
        self.compute_pressure_acceleration()
        self.update_fluid_velocity()
        self.update_fluid_position()

        self.rigid_solver.step()
        self.container.insert_object()
        self.rigid_solver.insert_rigid_object()
        self.renew_rigid_particle_state()

        self.enforce_domain_boundary_3D(self.container.material_fluid)

    _library_step_impl = _step

    def prepare(self):
        super().prepare()
        self.compute_pcisph_k()
