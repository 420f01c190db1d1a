"""DFSPH (reference: SPH/fluid_solvers/DFSPH.py): alpha factor, divergence-free and
constant-density solves."""
from .._native import T
from .base_solver import BaseSolver


class DFSPHSolver(BaseSolver):
    def __init__(self, container):
        super().__init__(container)
        self.m_max_iterations_v = 1000
        self.m_max_iterations = 1000
        self.m_eps = 1e-5
        self.max_error_V = 0.001
        self.max_error = 0.0001
        self.last_iterations_v = None
        self.last_iterations = None

    def compute_alpha(self):
        self._run(T.DFSPH_COMPUTE_ALPHA)

    def compute_density_derivative(self):
        self._run(T.DFSPH_COMPUTE_DENSITY_DERIVATIVE)

    def compute_density_star(self):
        self._run(T.DFSPH_COMPUTE_DENSITY_STAR)

    def compute_kappa_v(self):
        self._run(T.DFSPH_COMPUTE_KAPPA_V)

    def correct_divergence_step(self):
        self._run(T.DFSPH_CORRECT_DIVERGENCE_STEP)

    def compute_density_derivative_error(self) -> float:
        return self._run(T.DFSPH_COMPUTE_DENSITY_DERIVATIVE_ERROR)

    def compute_kappa(self):
        self._run(T.DFSPH_COMPUTE_KAPPA)

    def correct_density_error_step(self):
        self._run(T.DFSPH_CORRECT_DENSITY_ERROR_STEP)

    def compute_density_error(self) -> float:
        return self._run(T.DFSPH_COMPUTE_DENSITY_ERROR)

    def correct_divergence_error(self):
        num_itr = 0
        self.compute_density_derivative()
        err = 0.0
        while num_itr < 1 or num_itr < self.m_max_iterations_v:
            self.compute_kappa_v()
            self.correct_divergence_step()
            self.compute_density_derivative()
            err = self.compute_density_derivative_error()
            eta = self.max_error_V * self.density_0 / self.dt[None]
            num_itr += 1
            if err <= eta:
                break
        self.last_iterations_v = (num_itr, err)
        return num_itr

    def correct_density_error(self):
        self.compute_density_star()
        num_itr = 0
        err = 0.0
        while num_itr < 1 or num_itr < self.m_max_iterations:
            self.compute_kappa()
            self.correct_density_error_step()
            self.compute_density_star()
            err = self.compute_density_error()
            num_itr += 1
            if err <= self.max_error:
                break
        self.last_iterations = (num_itr, err)
        return num_itr

    def _step(self):
        self.compute_non_pressure_acceleration()
        self.update_fluid_velocity()
        self.correct_density_error()

        self.update_fluid_position()

        self.rigid_solver.step()
        self.container.insert_object()
        self.rigid_solver.insert_rigid_object()
        self.renew_rigid_particle_state()

        self.enforce_domain_boundary_3D(self.container.material_fluid)

        self.container.prepare_neighborhood_search()
        self.compute_density()
        self.compute_alpha()
        self.correct_divergence_error()

    _library_step_impl = _step

    def prepare(self):
        super().prepare()
        self.compute_density()
        self.compute_alpha()
