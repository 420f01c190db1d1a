"""Rigid-body stepper behind the reference's PyBulletSolver boundary (SURVEY.md 8(f3);
reference SPH/rigid_solver/bullet_solver.py:14-183).

Upstream hands the dynamic rigid bodies to PyBullet (third-party CPU physics, not installable
here).  What the fluid solvers need from that boundary is small: once per step, take the force and
torque the fluid kernels accumulated per object, advance each dynamic body, and write back centre of
mass, rotation, linear and angular velocity (bullet_solver.py:144-167); static bodies and scenes
without RigidBodies are no-ops (:40-42,145-146).  This class keeps that interface (same constructor,
`insert_rigid_object`, `step`, `total_time`, `present_rigid_object`) over a small built-in integrator:

  * free rigid-body dynamics, semi-implicit Euler, inertia tensor from the body's own particles about
    the base frame origin (upstream makes the same "centre of mass = base position" assumption, :12);
  * contacts with the six domain walls, inset like upstream's `create_boundary` (:50-71), resolved
    per step with an inelastic impulse at the deepest mesh vertex plus positional projection;
  * no body-body contacts and no friction (PyBullet has both): documented gap, bodies interact through
    the fluid only.

Results therefore cannot be pinned against PyBullet; tests check conservation laws and limits.
"""
from __future__ import annotations

import numpy as np


def _skew_exp(w, dt):
    """Rotation matrix exp([w dt]x) (Rodrigues)."""
    th = np.linalg.norm(w) * dt
    if th < 1e-12:
        return np.eye(3)
    k = w / np.linalg.norm(w)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _euler_xyz(rpy):
    """Rotation matrix of PyBullet's getQuaternionFromEuler([roll, pitch, yaw]): R = Rz(yaw) Ry(pitch) Rx(roll)."""
    (cr, sr), (cp, sp), (cy, sy) = [(np.cos(a), np.sin(a)) for a in rpy]
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


class _Body:
    def __init__(self, obj_id, mass, inertia_body, x, R, v, w, local_vertices):
        self.obj_id = obj_id
        self.mass = float(mass)
        self.inertia_body = inertia_body
        self.x, self.R, self.v, self.w = x, R, v, w
        self.local_vertices = local_vertices


class PyBulletSolver:
    def __init__(self, container, gravity=(0, -9.8, 0), dt=1e-3):
        self.container = container
        self.total_time = 0.0
        self.present_rigid_object = []
        assert container.dim == 3, "PyBulletSolver only supports 3D simulation currently"
        self.cfg = container.cfg
        self.rigid_bodies = self.cfg.get_rigid_bodies()
        self.rigid_blocks = self.cfg.get_rigid_blocks()
        self.dt = dt
        self.gravity = np.asarray(gravity, dtype=np.float64)
        self.bodies = {}
        num_rigid_bodies = len(self.rigid_bodies) + len(self.rigid_blocks)
        if num_rigid_bodies != 0:
            self.physicsClient = "builtin"   # upstream: a pybullet DIRECT client id
            self.create_boundary()
        else:
            self.physicsClient = None
            print("No rigid body in the scene, skip bullet solver initialization.")

    @property
    def is_noop(self):
        """True when step() has nothing to do: no rigid bodies at all, or only static ones."""
        return self.physicsClient is None or not any(b.get("isDynamic") for b in self.rigid_bodies)

    def create_boundary(self, thickness: float = 0.01):
        # the walls sit inside the fluid's domain box by this much (bullet_solver.py:50-55)
        c = self.container
        eps = c.padding + c.particle_diameter + c.domain_box_thickness
        self.wall_lo = np.array(c.domain_start, dtype=np.float64) + eps
        self.wall_hi = np.array(c.domain_end, dtype=np.float64) - eps

    # ------------------------------------------------------------------ insertion
    def insert_rigid_object(self):
        for rigid_body in self.rigid_bodies:
            self.init_rigid_body(rigid_body)
        for _ in self.rigid_blocks:
            raise NotImplementedError

    def init_rigid_body(self, rigid_body):
        obj = rigid_body["objectId"]
        if obj in self.present_rigid_object:
            return
        if rigid_body["entryTime"] > self.total_time:
            return
        c = self.container
        if rigid_body["isDynamic"]:
            translation = np.array(rigid_body["translation"], dtype=np.float64)
            angle = rigid_body["rotationAngle"] / 360 * (2 * np.pi)
            axis = np.array(rigid_body["rotationAxis"], dtype=np.float64)
            # upstream turns axis * angle into EULER angles (roll, pitch, yaw) — p.getQuaternionFromEuler(axis * angle),
            # bullet_solver.py:102-106 — not into an axis-angle rotation; the two agree for coordinate axes only
            R = _euler_xyz(axis * angle)
            velocity = np.array(rigid_body["velocity"], dtype=np.float64)
            # body-frame particles (the container inserted them unplaced, base_container.py:618-625)
            pts = np.asarray(rigid_body["voxelizedPoints"], dtype=np.float64)
            mp = float(rigid_body["density"]) * c.V0      # compute_rigid_body_mass (base_container.py:384-390)
            mass = mp * pts.shape[0]
            r2 = (pts ** 2).sum(1)
            inertia = mp * (np.eye(3) * r2.sum() - pts.T @ pts)
            self.bodies[obj] = _Body(obj, mass, inertia, translation.copy(), R, velocity.copy(), np.zeros(3),
                                     np.asarray(rigid_body["restPosition"], dtype=np.float64))
            c.rigid_body_original_centers_of_mass[obj] = np.zeros(3, dtype=np.float32)
            self._write_back(self.bodies[obj])
        self.present_rigid_object.append(obj)

    def _write_back(self, b):
        c = self.container
        c.rigid_body_centers_of_mass[b.obj_id] = b.x
        c.rigid_body_rotations[b.obj_id] = b.R
        c.rigid_body_velocities[b.obj_id] = b.v
        c.rigid_body_angular_velocities[b.obj_id] = b.w

    # ------------------------------------------------------------------ stepping
    def apply_force(self, container_idx, force):
        self._force[container_idx] = np.asarray(force, dtype=np.float64)

    def apply_torque(self, container_idx, torque):
        self._torque[container_idx] = np.asarray(torque, dtype=np.float64)

    def step(self):
        if self.physicsClient is None or not self.bodies:
            return
        c = self.container
        forces = c.rigid_body_forces.to_numpy()
        torques = c.rigid_body_torques.to_numpy()
        c.rigid_body_forces.fill(0.0)
        c.rigid_body_torques.fill(0.0)
        self._force, self._torque = {}, {}
        for obj in self.bodies:
            self.apply_force(obj, forces[obj])
            self.apply_torque(obj, torques[obj])
        for b in self.bodies.values():
            self._advance(b, self._force[b.obj_id], self._torque[b.obj_id])
            self._write_back(b)

    def _advance(self, b, force, torque):
        dt = self.dt
        I_world = b.R @ b.inertia_body @ b.R.T
        b.v = b.v + dt * (force / b.mass + self.gravity)
        b.w = b.w + dt * np.linalg.solve(I_world, torque - np.cross(b.w, I_world @ b.w))
        b.x = b.x + dt * b.v
        R = _skew_exp(b.w, dt) @ b.R
        u, _, vt = np.linalg.svd(R)          # keep R orthonormal
        b.R = u @ vt
        self._collide_walls(b)

    def _collide_walls(self, b):
        world = b.local_vertices @ b.R.T + b.x
        I_inv = np.linalg.inv(b.R @ b.inertia_body @ b.R.T)
        for axis in range(3):
            for sign, wall in ((+1.0, self.wall_lo[axis]), (-1.0, self.wall_hi[axis])):
                depth = sign * (wall - world[:, axis])          # > 0: vertex is through the wall
                k = int(np.argmax(depth))
                if depth[k] <= 0:
                    continue
                n = np.zeros(3)
                n[axis] = sign
                r = world[k] - b.x
                v_rel = b.v + np.cross(b.w, r)
                vn = float(v_rel @ n)
                if vn < 0:   # approaching: inelastic impulse along the wall normal
                    denom = 1.0 / b.mass + float(n @ np.cross(I_inv @ np.cross(r, n), r))
                    j = -vn / denom
                    b.v = b.v + j * n / b.mass
                    b.w = b.w + I_inv @ np.cross(r, j * n)
                b.x = b.x + depth[k] * n                         # project out of the wall
                world = world + depth[k] * n

    def get_rigid_body_states(self, container_idx):
        b = self.bodies[container_idx]
        return {"linear_velocity": b.v, "angular_velocity": b.w, "position": b.x, "rotation_matrix": b.R}
