"""Prescribed rigid-body dynamics for the coupling fixtures (test infrastructure).

PyBullet is not installable offline, so the rigid side of the reference's fluid<->rigid coupling is
replaced -- on BOTH sides of the comparison -- by this deliberately simple rule: a free body with the
unit inertia tensor the reference writes into its URDF (SPH/utils/urdf.py), semi-implicit Euler, no
contacts.  What the fixtures pin is everything around it: rigid particle insertion and mass, the
force / torque the fluid kernels accumulate per object, renew_rigid_particle_state, the Akinci volumes
of a moving body.  The applied wrench of every step is logged.
"""
import numpy as np


def rodrigues(w, dt):
    th = float(np.linalg.norm(w)) * dt
    if th < 1e-14:
        return np.eye(3)
    k = np.asarray(w, dtype=np.float64) / np.linalg.norm(w)
    K = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return np.eye(3) + np.sin(th) * K + (1.0 - np.cos(th)) * (K @ K)


class FreeBodyWorld:
    def __init__(self, dt=1e-3, gravity=(0.0, -9.81, 0.0)):
        self.dt = float(dt)
        self.gravity = np.asarray(gravity, dtype=np.float64)
        self.bodies = []
        self.log = []          # per step: {body: (force, torque)}

    def add_body(self, mass, position, rotation=None, velocity=(0.0, 0.0, 0.0)):
        self.bodies.append({"mass": float(mass), "x": np.array(position, dtype=np.float64),
                            "R": np.eye(3) if rotation is None else np.array(rotation, dtype=np.float64),
                            "v": np.array(velocity, dtype=np.float64), "w": np.zeros(3), "F": np.zeros(3), "T": np.zeros(3)})
        return len(self.bodies) - 1

    def apply_force(self, body, force):
        self.bodies[body]["F"] += np.asarray(force, dtype=np.float64)

    def apply_torque(self, body, torque):
        self.bodies[body]["T"] += np.asarray(torque, dtype=np.float64)

    def step(self):
        entry = {}
        for i, b in enumerate(self.bodies):
            entry[i] = (b["F"].copy(), b["T"].copy())
            if b["mass"] > 0.0:
                b["v"] = b["v"] + self.dt * (b["F"] / b["mass"] + self.gravity)
                b["w"] = b["w"] + self.dt * b["T"]          # unit inertia (the URDF's ixx = iyy = izz = 1)
                b["x"] = b["x"] + self.dt * b["v"]
                b["R"] = rodrigues(b["w"], self.dt) @ b["R"]
            b["F"] = np.zeros(3)
            b["T"] = np.zeros(3)
        self.log.append(entry)
