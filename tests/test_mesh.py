"""Mesh voxeliser: closed-form volumes, lattice anchoring, containers with mesh bodies
(CPU oracle engine; the particles then take the same kernels as the domain box)."""
import os

import numpy as np
import pytest

from helpers import by_uid, make_sim, oracle_library, scene
from sph_project_b200 import mesh as M


def write_box_obj(path, lo, hi):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    v = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    with open(path, "w") as f:
        for p in v:
            f.write("v %.9f %.9f %.9f\n" % tuple(p))
        for q in quads:
            f.write("f %d %d %d %d\n" % tuple(i + 1 for i in q))   # quads: exercises fan triangulation


def write_icosphere_obj(path, radius, center=(0, 0, 0), subdiv=3):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, float) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []
        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = (v[a] + v[b]) / 2
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    with open(path, "w") as fh:
        for p in v:
            fh.write("v %.9f %.9f %.9f\n" % tuple(np.asarray(center) + radius * p))
        for a, b, c in f:
            fh.write("f %d %d %d\n" % (a + 1, b + 1, c + 1))


def test_obj_loader_and_export_roundtrip(tmp_path):
    p = tmp_path / "box.obj"
    write_box_obj(p, [0, 0, 0], [1, 2, 3])
    m = M.load_obj(str(p))
    assert m.vertices.shape == (8, 3) and m.faces.shape == (12, 3)
    q = tmp_path / "again.obj"
    q.write_text(m.export("obj"))
    m2 = M.load_obj(str(q))
    assert np.allclose(m.vertices, m2.vertices) and np.array_equal(m.faces, m2.faces)


def test_box_voxel_count_and_anchoring(tmp_path):
    p = tmp_path / "box.obj"
    write_box_obj(p, [0.101, 0.101, 0.101], [0.299, 0.399, 0.199])
    pts = M.voxelize_filled_points(M.load_obj(str(p)), 0.02)
    # voxel centres are multiples of the pitch (trimesh anchors the voxel lattice at the origin)
    assert np.allclose(pts / 0.02, np.round(pts / 0.02), atol=1e-9)
    # round(x / pitch) over [0.101, 0.299] -> indices 5..15 (11), y 5..20 (16), z 5..10 (6)
    assert pts.shape[0] == 11 * 16 * 6
    assert np.isclose(pts.min(0), [0.10, 0.10, 0.10]).all() and np.isclose(pts.max(0), [0.30, 0.40, 0.20]).all()


def test_sphere_volume_and_inside(tmp_path):
    p = tmp_path / "sphere.obj"
    write_icosphere_obj(p, 0.2, center=(0.5, 0.5, 0.5))
    m = M.load_obj(str(p))
    pitch = 0.02
    pts = M.voxelize_filled_points(m, pitch)
    r = np.linalg.norm(pts - 0.5, axis=1)
    assert r.max() <= 0.2 + pitch * 0.87            # nothing further out than half a voxel diagonal
    vol = pts.shape[0] * pitch ** 3
    shell = 4 * np.pi * 0.2 ** 2 * pitch / 2        # surface voxels overshoot by about half a layer
    assert abs(vol - (4 / 3 * np.pi * 0.2 ** 3 + shell)) < 0.04 * vol
    # solid: no holes inside
    idx = set(map(tuple, np.round(pts / pitch).astype(int)))
    assert (25, 25, 25) in idx
    # point-in-mesh against the analytic sphere, away from the surface
    rng = np.random.default_rng(0)
    q = rng.uniform(0.25, 0.75, size=(4000, 3))
    d = np.linalg.norm(q - 0.5, axis=1)
    sure = np.abs(d - 0.2) > 0.004
    assert np.array_equal(M.points_inside(m, q)[sure], (d < 0.2)[sure])


def test_rotation_matrix():
    R = M.rotation_matrix(np.pi / 2, [0, 0, 1], [1, 0, 0])
    assert np.allclose(R @ np.array([2, 0, 0, 1]), [1, 1, 0, 1])
    assert np.allclose(M.rotation_matrix(0.0, [0, 1, 0], [3, 4, 5]), np.eye(4))


def test_static_rigid_and_fluid_body_scene(tmp_path):
    """A static mesh obstacle + a mesh-shaped fluid body go through the reference's scene schema."""
    box = tmp_path / "pillar.obj"
    write_box_obj(box, [-0.05, -0.1, -0.05], [0.05, 0.1, 0.05])
    ball = tmp_path / "ball.obj"
    write_icosphere_obj(ball, 0.08, subdiv=2)
    sc = scene("dfsph", dt=1e-3, domain_end=(0.8, 0.8, 0.8), block_start=(0.1, 0.1, 0.1), block_end=(0.3, 0.3, 0.3))
    sc["RigidBodies"] = [dict(objectId=1, geometryFile=str(box), translation=[0.5, 0.16, 0.4], rotationAxis=[0, 1, 0],
                              rotationAngle=30, scale=[1, 1, 1], velocity=[0, 0, 0], density=2200.0, color=[255, 255, 255],
                              isDynamic=False, entryTime=-1.0)]
    sc["FluidBodies"] = [dict(objectId=2, geometryFile=str(ball), translation=[0.5, 0.55, 0.4], rotationAxis=[0, 1, 0],
                              rotationAngle=0, scale=[1, 1, 1], velocity=[0, -1, 0], density=1000.0, color=[0, 0, 255],
                              entryTime=-1.0)]
    c, s = make_sim(sc, oracle_library())
    obj = by_uid(c, c.particle_object_ids)
    mat = by_uid(c, c.particle_materials)
    n_rigid, n_ball = int((obj == 1).sum()), int((obj == 2).sum())
    assert n_rigid == c.rigid_bodies[0]["particleNum"] > 300 and n_ball == c.fluid_bodies[0]["particleNum"] > 150
    assert (mat[obj == 1] == 2).all() and (mat[obj == 2] == 1).all()
    assert c.particle_num[None] == c.particle_max_num
    assert c.fluid_particle_num[None] == 1000 + n_ball
    # Akinci volumes were computed for the obstacle: an interior particle of a spacing-d lattice gets
    # V = d^3 = 1.25 V0 (V0 = 0.8 d^3), surface particles more
    V = by_uid(c, c.particle_rest_volumes)[obj == 1] / c.V0
    assert 1.24 < V.min() < 1.26 and V.max() < 3.0
    x0 = by_uid(c, c.particle_positions)[obj == 1].copy()
    s.step(20)
    x1 = by_uid(c, c.particle_positions)
    assert np.array_equal(x1[obj == 1], x0)                   # static body does not move
    assert x1[obj == 2, 1].mean() < 0.55 - 0.015             # the ball falls
    assert np.isfinite(x1).all()
