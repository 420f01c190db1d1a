"""Stand-in for the trimesh calls of the reference's base_container.py:611-717 (test infrastructure).

Delegates to the repository's own mesh module (sph_project_b200/mesh.py, loaded by file path so that the
reference's `SPH` package, not the repository's alias of the same name, stays importable): both sides of
the coupling fixtures then voxelise a body into the same particle set.  trimesh's own voxeliser is not
reproduced bit for bit (tests/test_mesh.py pins volumes instead).
"""
import importlib.util
import os

import numpy as np

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "sph_project_b200", "mesh.py")
_spec = importlib.util.spec_from_file_location("_repo_mesh", _path)
_mesh = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mesh)


class _Box:
    def __init__(self, bounds):
        self.bounds = bounds


class _Voxels:
    def __init__(self, mesh, pitch):
        self._mesh, self._pitch = mesh, pitch

    def fill(self):
        return self

    @property
    def points(self):
        return _mesh.voxelize_filled_points(self._mesh, self._pitch)


class Trimesh(_mesh.Mesh):
    def copy(self):
        return Trimesh(self.vertices.copy(), self.faces.copy())

    def voxelized(self, pitch):
        return _Voxels(_mesh.Mesh(self.vertices, self.faces), pitch)

    @property
    def bounding_box(self):
        return _Box(self.bounds)

    def contains(self, points):
        return _mesh.points_inside(self, np.asarray(points))


def load(path):
    m = _mesh.load_obj(path)
    return Trimesh(m.vertices, m.faces)


class repair:
    @staticmethod
    def fill_holes(mesh):
        return True


class transformations:
    rotation_matrix = staticmethod(_mesh.rotation_matrix)
