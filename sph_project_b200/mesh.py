"""Mesh bodies -> particles (reference base_container.py:611-717).

The reference uses trimesh (not installable here): rigid bodies are voxelised with
`mesh.voxelized(pitch).fill().points`, fluid bodies are the points of an `arange(min, max, pitch)`
lattice that lie inside the mesh (`mesh.contains`).  This module restates both with numpy + scipy:

  * `voxelize_surface`  trimesh's default "subdivide" voxeliser: split triangles until every edge is
                        shorter than pitch / 2, then a voxel is hit where round(vertex / pitch) lands;
  * `fill`              interior by `scipy.ndimage.binary_fill_holes`, like `VoxelGrid.fill()`;
  * points are voxel centres `index * pitch` (the lattice is anchored at the origin, as in trimesh);
  * `points_inside`     ray casting along +z with the even-odd rule.

Exact particle sets cannot be pinned against trimesh offline; the tests pin closed-form volumes
(cube, sphere) to a few percent and the lattice anchoring exactly.
"""
from __future__ import annotations

import numpy as np


class Mesh:
    """Triangle mesh (vertices [n,3] f64, faces [m,3] int) with the few trimesh calls upstream uses."""

    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)

    def copy(self):
        return Mesh(self.vertices.copy(), self.faces.copy())

    def apply_scale(self, scale):
        self.vertices = self.vertices * np.asarray(scale, dtype=np.float64)

    def apply_transform(self, matrix):
        m = np.asarray(matrix, dtype=np.float64)
        self.vertices = self.vertices @ m[:3, :3].T + m[:3, 3]

    @property
    def bounds(self):
        return self.vertices.min(0), self.vertices.max(0)

    def export(self, file_type="obj"):
        assert file_type == "obj"
        lines = [f"v {x:.8f} {y:.8f} {z:.8f}" for x, y, z in self.vertices]
        lines += [f"f {a + 1} {b + 1} {c + 1}" for a, b, c in self.faces]
        return "\n".join(lines) + "\n"


def load_obj(path) -> Mesh:
    """Wavefront OBJ: `v` and `f` records; polygons are fan-triangulated, `a/b/c` index forms accepted."""
    verts, faces = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                verts.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    k = int(tok.split("/")[0])
                    idx.append(k - 1 if k > 0 else len(verts) + k)
                for t in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[t], idx[t + 1]])
    if not verts or not faces:
        raise ValueError(f"{path}: no triangles found")
    return Mesh(verts, faces)


def rotation_matrix(angle, direction, point):
    """4x4 rotation by `angle` (radians) about the axis `direction` through `point`
    (trimesh.transformations.rotation_matrix)."""
    d = np.asarray(direction, dtype=np.float64)
    n = np.linalg.norm(d)
    m = np.eye(4)
    if n == 0.0 or angle == 0.0:
        return m
    d = d / n
    c, s = np.cos(angle), np.sin(angle)
    K = np.array([[0, -d[2], d[1]], [d[2], 0, -d[0]], [-d[1], d[0], 0]])
    R = c * np.eye(3) + s * K + (1 - c) * np.outer(d, d)
    p = np.asarray(point, dtype=np.float64)
    m[:3, :3] = R
    m[:3, 3] = p - R @ p
    return m


def _subdivide_to_size(vertices, faces, max_edge, max_iter=12):
    """Points covering the surface: split every triangle with an edge longer than max_edge into four
    until none is left; returns all vertices produced (duplicates are harmless for voxelisation)."""
    tri = vertices[faces]   # [m, 3, 3]
    done = []
    for _ in range(max_iter):
        if tri.shape[0] == 0:
            break
        e = np.stack([tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 1], tri[:, 0] - tri[:, 2]], 1)
        longest = np.sqrt((e ** 2).sum(-1)).max(1)
        small = longest <= max_edge
        done.append(tri[small].reshape(-1, 3))
        t = tri[~small]
        if t.shape[0] == 0:
            tri = t
            break
        a, b, c = t[:, 0], t[:, 1], t[:, 2]
        ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
        tri = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, c], 1),
                              np.stack([ab, bc, ca], 1)], 0)
    if tri.shape[0]:
        done.append(tri.reshape(-1, 3))
    return np.concatenate(done, 0) if done else np.zeros((0, 3))


def voxelize_surface(mesh: Mesh, pitch: float) -> np.ndarray:
    """Integer voxel indices hit by the surface (unique rows)."""
    pts = _subdivide_to_size(mesh.vertices, mesh.faces, pitch / 2.0)
    return np.unique(np.round(pts / pitch).astype(np.int64), axis=0)


def fill(indices: np.ndarray) -> np.ndarray:
    """Surface voxels + enclosed interior (VoxelGrid.fill(): scipy binary_fill_holes)."""
    from scipy import ndimage
    lo = indices.min(0)
    shape = indices.max(0) - lo + 1
    dense = np.zeros(shape, dtype=bool)
    dense[tuple((indices - lo).T)] = True
    filled = ndimage.binary_fill_holes(dense)
    return np.argwhere(filled) + lo


def voxelize_filled_points(mesh: Mesh, pitch: float) -> np.ndarray:
    """`mesh.voxelized(pitch).fill().points`: centres of the surface + interior voxels."""
    return fill(voxelize_surface(mesh, pitch)).astype(np.float64) * pitch


def points_inside(mesh: Mesh, points: np.ndarray) -> np.ndarray:
    """Even-odd ray cast along +z; points exactly on the surface may land on either side."""
    p = np.asarray(points, dtype=np.float64)
    tri = mesh.vertices[mesh.faces]
    inside = np.zeros(p.shape[0], dtype=bool)
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    # 2-D barycentric test in the xy plane per triangle, vectorised over points in blocks
    d = (b[:, 1] - c[:, 1]) * (a[:, 0] - c[:, 0]) + (c[:, 0] - b[:, 0]) * (a[:, 1] - c[:, 1])
    ok = np.abs(d) > 1e-300
    a, b, c, d = a[ok], b[ok], c[ok], d[ok]
    block = max(1, int(2e7 // max(len(d), 1)))
    for s in range(0, p.shape[0], block):
        q = p[s:s + block]
        px, py = q[:, 0][:, None], q[:, 1][:, None]
        l1 = ((b[:, 1] - c[:, 1]) * (px - c[:, 0]) + (c[:, 0] - b[:, 0]) * (py - c[:, 1])) / d
        l2 = ((c[:, 1] - a[:, 1]) * (px - c[:, 0]) + (a[:, 0] - c[:, 0]) * (py - c[:, 1])) / d
        l3 = 1.0 - l1 - l2
        hit = (l1 >= 0) & (l2 >= 0) & (l3 >= 0)
        z = l1 * a[:, 2] + l2 * b[:, 2] + l3 * c[:, 2]
        above = hit & (z > q[:, 2][:, None])
        inside[s:s + block] = (above.sum(1) % 2) == 1
    return inside


def _prepared_mesh(body, placed: bool) -> Mesh:
    mesh = load_obj(body["geometryFile"])
    mesh.apply_scale(body["scale"])
    if placed:
        # base_container.py:618-625 / :682-688: rotate about the vertex centroid, then translate
        angle = body["rotationAngle"] / 360 * 2 * 3.1415926
        mesh.apply_transform(rotation_matrix(angle, body["rotationAxis"], mesh.vertices.mean(axis=0)))
        mesh.vertices = mesh.vertices + np.array(body["translation"], dtype=np.float64)
    return mesh


def voxelize_rigid_body(rigid_body, pitch):
    """BaseContainer.load_rigid_body (base_container.py:611-643).  Static bodies are placed here;
    dynamic ones stay in their body frame (the rigid solver places them)."""
    mesh = _prepared_mesh(rigid_body, placed=not rigid_body["isDynamic"])
    backup = mesh.copy()
    rigid_body["mesh"] = backup
    rigid_body["restPosition"] = backup.vertices
    rigid_body["restCenterOfMass"] = np.array([0.0, 0.0, 0.0])
    points = voxelize_filled_points(mesh, pitch)
    print(f"rigid body {rigid_body['objectId']} num: {points.shape[0]}")
    return points


def voxelize_fluid_body(fluid_body, pitch, dim=3):
    """BaseContainer.load_fluid_body (base_container.py:676-717): lattice points inside the mesh."""
    mesh = _prepared_mesh(fluid_body, placed=True)
    lo, hi = mesh.bounds
    axes = [np.arange(lo[i], hi[i], pitch) for i in range(dim)]
    grid = np.array(np.meshgrid(*axes, sparse=False, indexing="ij"), dtype=np.float32)
    pts = grid.reshape(dim, -1).transpose()
    return pts[points_inside(mesh, pts)]
