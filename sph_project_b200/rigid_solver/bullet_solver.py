"""Rigid-body stepper boundary (reference: SPH/rigid_solver/bullet_solver.py:14-183).

Upstream drives PyBullet (third-party CPU physics, absent from this image and out of scope for the
CUDA hot path).  The boundary still has to exist because BaseSolver constructs it unconditionally
(base_solver.py:38); with no RigidBodies in the scene it is a no-op upstream as well
(bullet_solver.py:40-42,145-146).  Dynamic rigid bodies are a "next" row.
"""


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
        self.physicsClient = None
        if len(self.rigid_bodies) + len(self.rigid_blocks) == 0:
            print("No rigid body in the scene, skip bullet solver initialization.")
        elif any(b.get("isDynamic") for b in self.rigid_bodies):
            raise NotImplementedError("dynamic rigid bodies need a rigid-body stepper (PyBullet is not available)")

    @property
    def is_noop(self):
        return self.physicsClient is None

    def insert_rigid_object(self):
        for rigid_body in self.rigid_bodies:
            obj_id = rigid_body["objectId"]
            if obj_id in self.present_rigid_object or rigid_body["entryTime"] > self.total_time:
                continue
            self.present_rigid_object.append(obj_id)  # static bodies: particles are already in place
        for _ in self.rigid_blocks:
            raise NotImplementedError

    def step(self):
        if self.physicsClient is None:
            return
