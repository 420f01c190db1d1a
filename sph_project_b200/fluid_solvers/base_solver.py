"""Solver base class over the B200 SPH library.

Same public surface as the reference's BaseSolver (SPH/fluid_solvers/base_solver.py): every
upstream @ti.kernel is a method of the same name here and launches the corresponding hand-written
CUDA kernel through `sph_run_task`; `prepare()` / `step()` keep the upstream call order.  The
solver owns no particle state (upstream neither: base_solver.py:34-54); CG scratch for implicit
viscosity lives in the library and is exposed as the same `cg_*` attributes.
"""
from __future__ import annotations

import numpy as np

from .._native import F, S, T
from ..fields import ParticleField, ScalarField
from ..rigid_solver import PyBulletSolver


class BaseSolver:
    def __init__(self, container):
        self.container = container
        self.cfg = container.cfg
        self._engine = container.engine
        self.g = np.array(self.cfg.get_cfg("gravitation"))
        self.g_upper = self.cfg.get_cfg("gravitationUpper")
        if self.g_upper is None:
            self.g_upper = 10000.0  # a large number (base_solver.py:21-23)
        self.viscosity_method = self.cfg.get_cfg("viscosityMethod")
        if self.viscosity_method not in ("standard", "implicit"):
            # upstream raises on first use (base_solver.py:200); raising at construction is stricter
            raise NotImplementedError(f"viscosity method {self.viscosity_method} not implemented")
        self.viscosity = self.cfg.get_cfg("viscosity")
        self.viscosity_b = self.cfg.get_cfg("viscosity_b")
        if self.viscosity_b is None:
            self.viscosity_b = self.viscosity
        self.density_0 = self.cfg.get_cfg("density0")
        self.surface_tension = 0.01

        self.dt = ScalarField(self._engine, S.DT)
        self.dt[None] = self.cfg.get_cfg("timeStepSize")

        self.rigid_solver = PyBulletSolver(container, gravity=self.g, dt=self.dt[None])

        if self.viscosity_method == "implicit":
            eng, cap = self._engine, container.particle_max_num
            self.cg_p = ParticleField(eng, F.CG_P, cap)
            self.original_velocity = ParticleField(eng, F.ORIGINAL_VELOCITY, cap)
            self.cg_Ap = ParticleField(eng, F.CG_AP, cap)
            self.cg_x = ParticleField(eng, F.CG_X, cap)
            self.cg_b = ParticleField(eng, F.CG_B, cap)
            self.cg_alpha = ScalarField(eng, S.CG_ALPHA)
            self.cg_beta = ScalarField(eng, S.CG_BETA)
            self.cg_r = ParticleField(eng, F.CG_R, cap)
            self.cg_error = ScalarField(eng, S.CG_ERROR)
            self.cg_diagnol_ii_inv = ParticleField(eng, F.CG_DIAG_INV, cap, matrix=True)
            self.cg_tol = 1e-6
        self.last_stats = None

    def _run(self, task, iarg=0):
        return self._engine.run_task(task, iarg)

    # ---- upstream kernels, one launch each ----
    def compute_rigid_particle_volume(self):
        self._run(T.COMPUTE_RIGID_PARTICLE_VOLUME)

    def init_acceleration(self):
        self._run(T.INIT_ACCELERATION)

    def init_rigid_body_force_and_torque(self):
        self._run(T.INIT_RIGID_BODY_FORCE_AND_TORQUE)

    def compute_pressure_acceleration(self):
        self._run(T.COMPUTE_PRESSURE_ACCELERATION)

    # host-side equivalents of upstream's device functions (@ti.func, base_solver.py:56-103), f32 like the kernels;
    # for Python-side tasks over container.neighbor_lists() and for tests
    def kernel_W(self, R_mod):
        h = np.float32(self.container.dh)
        k = np.float32(8.0 / np.pi) / (h * h * h)
        q = np.asarray(R_mod, dtype=np.float32) / h
        inner = k * (np.float32(6.0) * q * q * q - np.float32(6.0) * q * q + np.float32(1.0))
        outer = k * np.float32(2.0) * np.power(np.maximum(np.float32(1.0) - q, np.float32(0.0)), np.float32(3.0))
        return np.where(q <= 0.5, inner, np.where(q <= 1.0, outer, np.float32(0.0))).astype(np.float32)

    def kernel_gradient(self, R):
        R = np.asarray(R, dtype=np.float32)
        h = np.float32(self.container.dh)
        k = np.float32(6.0) * np.float32(8.0 / np.pi) / (h * h * h)
        r = np.sqrt((R * R).sum(-1, dtype=np.float32))
        q = r / h
        safe = np.where(r > 1e-5, r, np.float32(1.0))
        grad_q = R / (safe * h)[..., None]
        mag = np.where(q <= 0.5, k * q * (np.float32(3.0) * q - np.float32(2.0)), -k * (np.float32(1.0) - q) ** 2)
        on = (r > 1e-5) & (q <= 1.0)
        return np.where(on[..., None], mag[..., None] * grad_q, np.float32(0.0)).astype(np.float32)

    def compute_non_pressure_acceleration(self):
        # gravity, surface tension and viscosity (base_solver.py:190-200)
        self.compute_gravity_acceleration()
        self.compute_surface_tension_acceleration()
        if self.viscosity_method == "standard":
            self.compute_viscosity_acceleration_standard()
        elif self.viscosity_method == "implicit":
            self.implicit_viscosity_solve()
        else:
            raise NotImplementedError(f"viscosity method {self.viscosity_method} not implemented")

    def compute_gravity_acceleration(self):
        self._run(T.COMPUTE_GRAVITY_ACCELERATION)

    def compute_surface_tension_acceleration(self):
        self._run(T.COMPUTE_SURFACE_TENSION_ACCELERATION)

    def compute_viscosity_acceleration_standard(self):
        self._run(T.COMPUTE_VISCOSITY_ACCELERATION_STANDARD)

    # implicit viscosity (base_solver.py:281-517)
    def prepare_conjugate_gradient_solver1(self):
        self._run(T.CG_PREPARE1)

    def prepare_conjugate_gradient_solver2(self):
        self._run(T.CG_PREPARE2)

    def compute_Ap(self):
        self._run(T.CG_COMPUTE_AP)

    def compute_cg_alpha(self):
        self._run(T.CG_COMPUTE_ALPHA)

    def update_cg_x(self):
        self._run(T.CG_UPDATE_X)

    def update_cg_r_and_beta(self):
        return self._run(T.CG_UPDATE_R_AND_BETA)

    def update_p(self):
        self._run(T.CG_UPDATE_P)

    def prepare_guess(self):
        self._run(T.CG_PREPARE_GUESS)

    def conjugate_gradient_loop(self):
        tol, num_itr = 1000.0, 0
        while tol > self.cg_tol and num_itr < 1000:
            self.compute_Ap()
            self.compute_cg_alpha()
            self.update_cg_x()
            tol = self.update_cg_r_and_beta()  # returns cg_error: one device->host read per iteration
            self.update_p()
            num_itr += 1
        self.last_cg = (num_itr, tol)
        return num_itr

    def viscosity_update_velocity(self):
        self._run(T.VISCOSITY_UPDATE_VELOCITY)

    def copy_back_original_velocity(self):
        self._run(T.COPY_BACK_ORIGINAL_VELOCITY)

    def implicit_viscosity_solve(self):
        self.prepare_conjugate_gradient_solver1()
        self.compute_Ap()
        self.prepare_conjugate_gradient_solver2()
        self.conjugate_gradient_loop()
        self.viscosity_update_velocity()
        self.compute_viscosity_acceleration_standard()  # accelerations from the solved velocities
        self.copy_back_original_velocity()
        self.prepare_guess()

    def compute_density(self):
        self._run(T.COMPUTE_DENSITY)

    def enforce_domain_boundary_3D(self, particle_type: int):
        self._run(T.ENFORCE_DOMAIN_BOUNDARY_3D, particle_type)

    def enforce_domain_boundary_2D(self, particle_type: int):
        raise NotImplementedError("2-D is unreachable upstream (bullet_solver.py:19) and not built")

    def enforce_domain_boundary(self, particle_type: int):
        self.enforce_domain_boundary_3D(particle_type)

    def _renew_rigid_particle_state(self):
        self._run(T.RENEW_RIGID_PARTICLE_STATE)

    def renew_rigid_particle_state(self):
        self._renew_rigid_particle_state()
        if self.cfg.get_cfg("exportObj"):   # keep the export meshes in step (base_solver.py:634-640)
            c = self.container
            for obj_i in range(c.object_num[None]):
                if c.rigid_body_is_dynamic[obj_i] and c.object_materials[obj_i] == c.material_rigid:
                    body = c.object_collection.get(obj_i)
                    if not isinstance(body, dict) or "mesh" not in body:
                        continue
                    rot = np.asarray(c.rigid_body_rotations[obj_i], dtype=np.float64)
                    com = np.asarray(c.rigid_body_centers_of_mass[obj_i], dtype=np.float64)
                    body["mesh"].vertices = (rot @ (body["restPosition"] - body["restCenterOfMass"]).T).T + com

    def update_fluid_velocity(self):
        self._run(T.UPDATE_FLUID_VELOCITY)

    def update_fluid_position(self):
        self._run(T.UPDATE_FLUID_POSITION)

    def prepare_emitter(self):
        self._run(T.PREPARE_EMITTER)

    def init_object_id(self):
        self._run(T.INIT_OBJECT_ID)

    # ---- orchestration (base_solver.py:683-696) ----
    def prepare(self):
        self.init_object_id()
        self.container.insert_object()
        self.prepare_emitter()
        self.rigid_solver.insert_rigid_object()
        self.renew_rigid_particle_state()
        self.container.prepare_neighborhood_search()
        self.compute_rigid_particle_volume()

    def _pending_objects(self) -> bool:
        c = self.container
        objs = list(c.fluid_blocks) + list(c.fluid_bodies) + list(c.rigid_bodies)
        return any(o["objectId"] not in c.present_object for o in objs)

    def _native_step_ok(self) -> bool:
        """The whole step can run inside the library when the Python hooks in the middle of
        `_step` (rigid_solver.step, insert_object) have nothing to do and `_step` is not overridden."""
        return (self.rigid_solver.is_noop and not self._pending_objects()
                and type(self)._step is type(self)._library_step_impl)

    _library_step_impl = None  # set by subclasses to their own `_step`

    def step(self, n_steps: int = 1):
        if self._native_step_ok():
            self.last_stats = self._engine.step(n_steps)
            dt = self.dt[None]
            self.container.total_time += dt * n_steps
            self.rigid_solver.total_time += dt * n_steps
            return self.last_stats
        for _ in range(n_steps):
            self._step()
            self.container.total_time += self.dt[None]
            self.rigid_solver.total_time += self.dt[None]
            self.compute_rigid_particle_volume()
        return None
