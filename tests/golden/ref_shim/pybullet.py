"""Stand-in for the pybullet calls of the reference's SPH/rigid_solver/bullet_solver.py (test infrastructure).

Scenes without RigidBodies never reach it (bullet_solver.py:40-42,144-146).  For the coupling fixtures the
dynamics are free_body.FreeBodyWorld -- NOT Bullet: no contacts, unit inertia; see that module.
"""
import re

import numpy as np

from free_body import FreeBodyWorld

DIRECT = 0
GEOM_BOX = 3
WORLD_FRAME = 1

world = None
_static = []


def connect(mode):
    global world
    world = FreeBodyWorld()
    return 0


def setAdditionalSearchPath(path):
    pass


def setTimeStep(dt):
    world.dt = float(dt)


def setGravity(gx, gy, gz):
    world.gravity = np.array([gx, gy, gz], dtype=np.float64)


def getQuaternionFromEuler(e):
    r, p, y = (float(v) * 0.5 for v in e)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def getMatrixFromQuaternion(q):
    x, y, z, w = q
    return (1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y))


def _quat_from_matrix(R):
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    if w > 1e-8:
        return ((R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w)
    raise NotImplementedError("half-turn orientations are not needed by the fixtures")


def createCollisionShape(kind, halfExtents=None):
    return len(_static)


def createMultiBody(baseMass=0, baseCollisionShapeIndex=-1, basePosition=(0, 0, 0)):
    _static.append(tuple(basePosition))    # the six domain walls: no contacts in the prescribed dynamics
    return -1


def loadURDF(path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1)):
    mass = float(re.search(r'<mass value="([^"]+)"', open(path).read()).group(1))
    R = np.array(getMatrixFromQuaternion(baseOrientation)).reshape(3, 3)
    return world.add_body(mass, basePosition, R)


def resetBaseVelocity(body, velocity):
    world.bodies[body]["v"] = np.array(velocity, dtype=np.float64)


def changeDynamics(body, link, mass=None):
    if mass is not None:
        world.bodies[body]["mass"] = float(mass)


def getBasePositionAndOrientation(body):
    b = world.bodies[body]
    return tuple(b["x"]), _quat_from_matrix(b["R"])


def getBaseVelocity(body):
    b = world.bodies[body]
    return tuple(b["v"]), tuple(b["w"])


def applyExternalForce(body, link, forceObj=None, posObj=None, flags=None):
    world.apply_force(body, forceObj.to_numpy() if hasattr(forceObj, "to_numpy") else forceObj)


def applyExternalTorque(body, link, torqueObj=None, flags=None):
    world.apply_torque(body, torqueObj.to_numpy() if hasattr(torqueObj, "to_numpy") else torqueObj)


def stepSimulation():
    world.step()
