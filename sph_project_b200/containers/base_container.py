"""Particle container over the B200 SPH library.

Mirrors the surface of the reference's BaseContainer (SPH/containers/base_container.py) —
constructor, attribute names, `insert_object`, `add_particles`, `add_cube`, `add_box`,
`prepare_neighborhood_search`, `for_all_neighbors`, `copy_to_vis_buffer`, `dump` and the
particle-count helpers — but owns no Taichi fields: every `particle_*` attribute is a view of
device memory held by the C-ABI library (include/sph_b200.h), and the uniform grid, counting
sort and neighbour walks run as sm_100a CUDA kernels.

Scene -> particles stays on the host in numpy exactly as upstream does it (f64 `np.arange`,
`meshgrid(indexing='ij')`, cast to f32), because those lattices define the initial condition.
"""
from __future__ import annotations

from functools import reduce

import numpy as np

from .. import _native as nat
from .._native import F, S
from ..fields import HostField, HostScalar, ObjectTable, ParticleField, ScalarField, WrenchTable
from ..utils import SimConfig

_METHOD_IDS = {"wcsph": nat.METHOD_WCSPH, "pcisph": nat.METHOD_PCISPH, "dfsph": nat.METHOD_DFSPH}


def _lattice_axes(lower_corner, extent, space, dim):
    return [np.arange(lower_corner[i], lower_corner[i] + extent[i], space) for i in range(dim)]


def _lattice(lower_corner, extent, space, dim):
    """Regular lattice like base_container.py:770-781: f64 arange per axis, ij-meshgrid, f32."""
    axes = _lattice_axes(lower_corner, extent, space, dim)
    grid = np.array(np.meshgrid(*axes, sparse=False, indexing="ij"), dtype=np.float32)
    return grid.reshape(dim, -1).transpose()


# ---- the same lattices cut along z (Z-slabs, not upstream): a rank builds the rows of its own cell layers only.
# The ij-meshgrid flattens x slowest / z fastest, so the row number of lattice point (ix, iy, iz) is known without
# building the lattice, and the box-shell test of base_container.py:830-835 separates into one mask per axis.
# tests/test_slab_host.py checks every function below against the full lattice filtered afterwards, bit for bit.

def _lattice_rows(axes, kz):
    """Rows of the 3-D lattice whose z index is in `kz` (ascending), in lattice order, and their row numbers in the
    full lattice."""
    nx, ny, nz = (len(a) for a in axes)
    grid = np.array(np.meshgrid(axes[0], axes[1], axes[2][kz], sparse=False, indexing="ij"), dtype=np.float32)
    rows = (np.arange(nx * ny, dtype=np.int64)[:, None] * nz + kz[None, :]).reshape(-1)
    return grid.reshape(3, -1).transpose(), rows


def _shell_masks(axes, lower_corner, cube_size, thickness):
    """Per axis: which lattice coordinates lie inside a wall of the hollow box (the test is applied to the f32
    coordinate, as `_box_shell` applies it to the f32 positions)."""
    masks = []
    for i, a in enumerate(axes):
        a32 = a.astype(np.float32)
        masks.append((a32 <= lower_corner[i] + thickness) | (a32 >= lower_corner[i] + cube_size[i] - thickness))
    return masks


def _shell_z_counts(masks):
    """Shell particles per z index: a whole x-y plane inside the floor / lid, the rim of the plane elsewhere."""
    mx, my, mz = masks
    rim = int((mx[:, None] | my[None, :]).sum())
    return np.where(mz, mx.size * my.size, rim).astype(np.int64)


def _shell_rows(axes, masks, kz):
    """Rows of the hollow box on the z indices `kz` (ascending), in lattice order, and their row numbers among the
    shell rows of the whole box."""
    nx, ny, nz = (len(a) for a in axes)
    mx, my, mz = masks
    wall = mx[:, None] | my[None, :]                       # columns inside an x or y wall keep every z
    per_column = np.where(wall, nz, int(mz.sum())).astype(np.int64).reshape(-1)
    column_base = (np.cumsum(per_column) - per_column).reshape(nx, ny)
    rank_z = np.cumsum(mz) - 1                             # place of z index k among the kept z of an open column
    keep = wall[:, :, None] | mz[kz][None, None, :]
    rows = column_base[:, :, None] + np.where(wall[:, :, None], kz[None, None, :], rank_z[kz][None, None, :])
    grid = np.array(np.meshgrid(axes[0], axes[1], axes[2][kz], sparse=False, indexing="ij"), dtype=np.float32)
    return grid[:, keep].transpose(), rows[keep]


class _NoOpScan:
    """Stands where upstream keeps `ti.algorithms.PrefixSumExecutor`: the scan is part of the library's sort."""

    def run(self, field):
        pass


class _GridCountsView:
    """Read-only stand-in for the `grid_num_particles` Taichi field (to_numpy / [i])."""

    def __init__(self, engine):
        self._engine = engine

    def to_numpy(self):
        return self._engine.get_grid_num_particles()

    def __getitem__(self, i):
        return self.to_numpy()[i]


class BaseContainer:
    # which solver's extra fields the library allocates; subclasses override
    _method = "wcsph"

    def __init__(self, config: SimConfig, GGUI=False, *, engine_library=None, device=0, slab=False):
        self.cfg = config
        self.GGUI = GGUI
        self.total_time = 0.0

        self.domain_start = np.array(self.cfg.get_cfg("domainStart"))
        assert self.domain_start[1] >= 0.0, "domain start y should be greater than 0"
        self.domain_end = np.array(self.cfg.get_cfg("domainEnd"))
        self.domain_size = self.domain_end - self.domain_start
        self.dim = len(self.domain_size)
        # the reference cannot build a 2-D solver either (rigid_solver/bullet_solver.py:19)
        assert self.dim == 3, "only 3D simulation is supported"
        print(f"Dimension: {self.dim}")

        self.material_rigid = nat.MATERIAL_RIGID
        self.material_fluid = nat.MATERIAL_FLUID

        self.dx = self.cfg.get_cfg("particleRadius")  # particle radius
        self.particle_diameter = 2 * self.dx
        self.dh = self.dx * 4.0  # support radius (3-D)
        if self.cfg.get_cfg("supportRadius"):
            self.dh = self.cfg.get_cfg("supportRadius")
        self.particle_spacing = self.particle_diameter
        if self.cfg.get_cfg("particleSpacing"):
            self.particle_spacing = self.cfg.get_cfg("particleSpacing")
        self.V0 = 0.8 * self.particle_diameter ** self.dim

        self.max_num_object = nat.MAX_OBJECTS

        # uniform grid: cell edge = support radius
        self.grid_size = self.dh
        self.grid_num = np.ceil(self.domain_size / self.grid_size).astype(int)
        print("grid size: ", self.grid_num)
        self.padding = self.grid_size

        self.add_domain_box = self.cfg.get_cfg("addDomainBox")
        if self.add_domain_box:
            self.domain_box_start = [self.domain_start[i] + self.padding for i in range(self.dim)]
            self.domain_box_size = [self.domain_size[i] - 2 * self.padding for i in range(self.dim)]
            self.domain_box_thickness = 0.03
        else:
            self.domain_box_thickness = 0.0

        self.object_collection = dict()
        self.object_id_rigid_body = set()
        self.object_id_fluid_body = set()
        self.present_object = []

        # ---- count particles (base_container.py:74-122) ----
        fluid_particle_num = 0
        rigid_body_particle_num = 0

        self.fluid_bodies = self.cfg.get_fluid_bodies()
        for fluid_body in self.fluid_bodies:
            points = self.load_fluid_body(fluid_body, pitch=self.particle_spacing)
            fluid_body["particleNum"] = points.shape[0]
            fluid_body["voxelizedPoints"] = points
            fluid_particle_num += points.shape[0]

        self.fluid_blocks = self.cfg.get_fluid_blocks()
        for fluid in self.fluid_blocks:
            particle_num = self.compute_cube_particle_num(fluid["start"], fluid["end"], space=self.particle_spacing)
            fluid["particleNum"] = particle_num
            fluid_particle_num += particle_num

        num_fluid_object = len(self.fluid_blocks) + len(self.fluid_bodies)

        self.rigid_bodies = self.cfg.get_rigid_bodies()
        for rigid_body in self.rigid_bodies:
            points = self.load_rigid_body(rigid_body, pitch=self.particle_spacing)
            rigid_body["particleNum"] = points.shape[0]
            rigid_body["voxelizedPoints"] = points
            rigid_body_particle_num += points.shape[0]

        self.rigid_blocks = self.cfg.get_rigid_blocks()
        for _ in self.rigid_blocks:
            raise NotImplementedError  # same as upstream (base_container.py:105-106)

        num_rigid_object = len(self.rigid_blocks) + len(self.rigid_bodies)
        print(f"Number of rigid bodies and rigid blocks: {num_rigid_object}")

        self.rigid_body_particle_num_total = rigid_body_particle_num
        box_particles = (
            self.compute_box_particle_num(self.domain_box_start, self.domain_box_size, space=self.particle_spacing,
                                          thickness=self.domain_box_thickness)
            if self.add_domain_box else 0
        )
        self.particle_max_num = int(fluid_particle_num + rigid_body_particle_num + box_particles)
        print(f"Fluid particle num: {fluid_particle_num}, Rigid body particle num: {rigid_body_particle_num}")

        # ---- Z-slab sharding (multi-GPU; not upstream): this rank keeps the particles of its slab ----
        self.global_particle_num = self.particle_max_num
        self.slab = None
        self._next_uid = 0
        if slab:
            self.slab = self._make_slab_context(*slab)
            self.particle_max_num = self.slab.capacity(self._layer_counts)

        # ---- device state: one library handle instead of ~35 Taichi fields (:129-190) ----
        self._engine = nat.Engine(self._make_params(device, bool(slab)), lib=engine_library)
        eng, cap = self._engine, self.particle_max_num
        if self.slab is not None:
            from ..slab import broadcast_bytes
            uid = self._engine.slab_unique_id() if self.slab.rank == 0 else None
            uid = broadcast_bytes(uid, 128, src=0)
            self._engine.slab_init(self.slab.rank, self.slab.world, uid, self.slab.z_lo, self.slab.z_hi,
                                   self.global_particle_num)
            self._connect_slab_peers()

        self.particle_num = ScalarField(eng, S.PARTICLE_NUM, int)
        self.fluid_particle_num = ScalarField(eng, S.FLUID_PARTICLE_NUM, int)

        def field(fid, matrix=False):
            return ParticleField(eng, fid, cap, matrix)

        self.particle_object_ids = field(F.OBJECT_ID)
        self.particle_positions = field(F.POSITION)
        self.particle_velocities = field(F.VELOCITY)
        self.particle_accelerations = field(F.ACCELERATION)
        self.particle_rest_volumes = field(F.REST_VOLUME)
        self.particle_masses = field(F.MASS)
        self.particle_densities = field(F.DENSITY)
        self.particle_pressures = field(F.PRESSURE)
        self.particle_materials = field(F.MATERIAL)
        self.particle_colors = field(F.COLOR)
        self.particle_is_dynamic = field(F.IS_DYNAMIC)
        self.rigid_particle_original_positions = field(F.ORIGINAL_POSITION)
        self.grid_ids = field(F.GRID_ID)
        self.particle_uids = field(F.UID)  # not upstream: insertion index carried through every sort

        self.object_materials = ObjectTable((), np.int32, self._push_object)
        self.prefix_sum_executor = _NoOpScan()
        self.object_num = HostScalar(num_fluid_object + num_rigid_object + (1 if self.add_domain_box else 0))
        self.rigid_body_is_dynamic = ObjectTable((), np.int32, self._push_object)
        self.rigid_body_original_centers_of_mass = ObjectTable((3,), np.float32, self._push_rigid)
        self.rigid_body_masses = ObjectTable((), np.float32)
        self.rigid_body_centers_of_mass = ObjectTable((3,), np.float32, self._push_rigid)
        self.rigid_body_rotations = ObjectTable((3, 3), np.float32, self._push_rigid)
        self.rigid_body_forces = WrenchTable(eng, 0)
        self.rigid_body_torques = WrenchTable(eng, 1, self.rigid_body_forces.state)
        self.rigid_body_velocities = ObjectTable((3,), np.float32, self._push_rigid)
        self.rigid_body_angular_velocities = ObjectTable((3,), np.float32, self._push_rigid)
        self.rigid_body_particle_num = ObjectTable((), np.int32)
        self.object_visibility = ObjectTable((), np.int32)

        self.x_vis_buffer = None
        if self.GGUI:
            self.x_vis_buffer = HostField((cap, self.dim))
            self.color_vis_buffer = HostField((cap, 3))

        if self.add_domain_box:
            box_id = self.object_num[None] - 1  # the last object id goes to the domain box
            self.add_box(object_id=box_id, lower_corner=self.domain_box_start, cube_size=self.domain_box_size,
                         thickness=self.domain_box_thickness, material=self.material_rigid, is_dynamic=False,
                         space=self.particle_spacing, color=(127, 127, 127))
            self.object_visibility[box_id] = 0
            self.object_materials[box_id] = self.material_rigid
            self.rigid_body_is_dynamic[box_id] = 0
            self.rigid_body_velocities[box_id] = [0.0 for _ in range(self.dim)]
            self.object_collection[box_id] = 0  # dummy

    # ------------------------------------------------------------------ library plumbing
    def _make_params(self, device, slab):
        cfg = self.cfg
        p = nat.SphParams()
        p.abi_version = nat.ABI_VERSION
        p.dim = self.dim
        p.method = _METHOD_IDS[self._method]
        visc_method = cfg.get_cfg("viscosityMethod")
        p.visc_method = nat.VISC_IMPLICIT if visc_method == "implicit" else nat.VISC_STANDARD
        p.max_particles = self.particle_max_num
        for i in range(3):
            p.grid_num[i] = int(self.grid_num[i])
            p.domain_size[i] = float(self.domain_size[i])
        p.dx, p.dh, p.V0 = float(self.dx), float(self.dh), float(self.V0)
        p.density0 = float(cfg.get_cfg("density0"))
        p.dt = float(cfg.get_cfg("timeStepSize"))
        g = cfg.get_cfg("gravitation")
        for i in range(3):
            p.gravity[i] = float(g[i])
        g_upper = cfg.get_cfg("gravitationUpper")
        p.g_upper = 10000.0 if g_upper is None else float(g_upper)
        p.viscosity = float(cfg.get_cfg("viscosity"))
        visc_b = cfg.get_cfg("viscosity_b")
        p.viscosity_b = p.viscosity if visc_b is None else float(visc_b)
        p.surface_tension = 0.01
        p.padding = float(self.padding)
        p.device = int(device)
        p.flags = nat.FLAG_SLAB if slab else 0
        return p

    @property
    def engine(self) -> nat.Engine:
        return self._engine

    def _block_lattice(self, fluid):
        offset = np.array(fluid["translation"])
        start = np.array(fluid["start"]) + offset
        end = np.array(fluid["end"]) + offset
        return start, (end - start) * np.array(fluid["scale"])

    def _scene_layers(self, nz):
        """Particles per cell layer of everything the scene will ever insert, as (counts, is_fluid) per object, for the
        slab split.  Box and blocks are counted from their lattice axes (no rank builds the whole scene); mesh bodies
        from their voxelised points."""
        from ..slab import cell_layer
        parts = []

        def by_layer(z_values, per_value):
            return np.bincount(cell_layer(z_values, self.dh, nz), weights=per_value, minlength=nz).astype(np.int64)

        if self.add_domain_box:
            axes = _lattice_axes(self.domain_box_start, self.domain_box_size, self.particle_spacing, self.dim)
            masks = _shell_masks(axes, self.domain_box_start, self.domain_box_size, self.domain_box_thickness)
            parts.append((by_layer(axes[2], _shell_z_counts(masks)), False))
        for fluid in self.fluid_blocks:
            axes = _lattice_axes(*self._block_lattice(fluid), self.particle_spacing, self.dim)
            parts.append((by_layer(axes[2], np.full(len(axes[2]), len(axes[0]) * len(axes[1]))), True))
        for bodies, is_fluid in ((self.fluid_bodies, True), (self.rigid_bodies, False)):
            for body in bodies:
                pts = np.asarray(body["voxelizedPoints"], dtype=np.float32)
                if len(pts):
                    parts.append((by_layer(pts[:, 2], np.ones(len(pts))), is_fluid))
        return parts

    # what a boundary particle costs relative to a fluid particle: it is sorted and gathered like any other, but the
    # neighbour sweeps (> 90 % of a step) only work on fluid rows.  Measured: with slabs balanced by plain particle
    # count, the thicker front wall of the tank left one of two ranks 10 % more fluid and the other waiting for it
    # twice per solver iteration (profiles/r02_bench_2gpu_*.json).
    SLAB_BOUNDARY_WEIGHT = 0.05

    def _make_slab_context(self, rank, world):
        from ..slab import SlabContext, balanced_ranges
        if self.dim != 3:
            raise ValueError("Z-slabs need a 3-D scene")
        nz = int(self.grid_num[2])
        counts = np.zeros(nz, dtype=np.int64)
        work = np.zeros(nz, dtype=np.float64)
        for layer, is_fluid in self._scene_layers(nz):
            counts += layer
            work += layer * (1.0 if is_fluid else self.SLAB_BOUNDARY_WEIGHT)
        self._layer_counts = counts   # capacities follow the true counts, the cut follows the work
        return SlabContext(rank=int(rank), world=int(world), dh=float(self.dh), nz=nz,
                           ranges=balanced_ranges(np.rint(work * 100).astype(np.int64), int(world)))

    def _connect_slab_peers(self):
        """Exchange the CUDA IPC handles that let the solver loops read ghosts from the neighbours' memory over NVLink
        (include/sph_b200.h, sph_slab_peer_*).  Optional: any failure leaves the NCCL path in place on every rank."""
        import os
        import torch.distributed as dist
        ok, blob = 1, b""
        if os.environ.get("SPH_B200_NO_PEER") == "1" or dist.get_backend() != "nccl":
            ok = 0
        else:
            try:
                blob = self._engine.slab_peer_export()
            except nat.SphError:
                ok = 0
        blobs = [None] * self.slab.world
        dist.all_gather_object(blobs, (ok, blob))
        if not all(o for o, _ in blobs):
            return
        for r, (_, b) in enumerate(blobs):
            if r != self.slab.rank:
                self._engine.slab_peer_import(r, b)

    def owned_mask(self):
        """Live particles this rank owns (everything unless the container is a Z-slab: then not the ghosts)."""
        n = self.particle_num[None]
        mask = np.ones(n, dtype=bool)
        if self.slab is not None:
            info = self._engine.slab_info()
            if info.own_end > info.own_begin or info.n_ghost:
                mask[:] = False
                mask[info.own_begin:info.own_end] = True
        return mask

    def _push_object(self, obj_id):
        self._engine.set_object(obj_id, int(self.object_materials[obj_id]), int(self.rigid_body_is_dynamic[obj_id]))

    def _push_rigid(self, obj_id):
        self._engine.set_rigid_state(
            obj_id, self.rigid_body_original_centers_of_mass[obj_id], self.rigid_body_centers_of_mass[obj_id],
            self.rigid_body_rotations[obj_id], self.rigid_body_velocities[obj_id],
            self.rigid_body_angular_velocities[obj_id])

    # ------------------------------------------------------------------ scene -> particles
    def insert_object(self):
        """Insert every object whose entryTime has come (base_container.py:212-341)."""
        for fluid in self.fluid_blocks:
            obj_id = fluid["objectId"]
            if obj_id in self.present_object or fluid["entryTime"] > self.total_time:
                continue
            offset = np.array(fluid["translation"])
            start = np.array(fluid["start"]) + offset
            end = np.array(fluid["end"]) + offset
            scale = np.array(fluid["scale"])
            self.object_id_fluid_body.add(obj_id)
            self.object_visibility[obj_id] = fluid.get("visible", 1)
            self.object_materials[obj_id] = self.material_fluid
            self.object_collection[obj_id] = fluid
            self.add_cube(object_id=obj_id, lower_corner=start, cube_size=(end - start) * scale,
                          velocity=fluid["velocity"], density=fluid["density"], is_dynamic=1, color=fluid["color"],
                          material=self.material_fluid, space=self.particle_spacing)
            self.present_object.append(obj_id)

        for fluid_body in self.fluid_bodies:
            obj_id = fluid_body["objectId"]
            if obj_id in self.present_object or fluid_body["entryTime"] > self.total_time:
                continue
            n = fluid_body["particleNum"]
            self.object_visibility[obj_id] = fluid_body.get("visible", 1)
            self.object_materials[obj_id] = self.material_fluid
            self.object_id_fluid_body.add(obj_id)
            self.object_collection[obj_id] = fluid_body
            self._add_uniform(obj_id, fluid_body["voxelizedPoints"], fluid_body["velocity"], fluid_body["density"],
                              self.material_fluid, 1, fluid_body["color"])
            self.present_object.append(obj_id)
            self.fluid_particle_num[None] += n

        for rigid_body in self.rigid_bodies:
            obj_id = rigid_body["objectId"]
            if obj_id in self.present_object or rigid_body["entryTime"] > self.total_time:
                continue
            self.object_id_rigid_body.add(obj_id)
            n = rigid_body["particleNum"]
            self.rigid_body_particle_num[obj_id] = n
            is_dynamic = rigid_body["isDynamic"]
            velocity = np.array(rigid_body["velocity"], dtype=np.float32) if is_dynamic else np.zeros(self.dim, np.float32)
            self.object_visibility[obj_id] = rigid_body.get("visible", 1)
            self.object_materials[obj_id] = self.material_rigid
            self.object_collection[obj_id] = rigid_body
            self._add_uniform(obj_id, rigid_body["voxelizedPoints"], velocity, rigid_body["density"],
                              self.material_rigid, int(bool(is_dynamic)), rigid_body["color"])
            self.rigid_body_is_dynamic[obj_id] = int(bool(is_dynamic))
            self.rigid_body_velocities[obj_id] = velocity
            if is_dynamic:
                self.rigid_body_masses[obj_id] = self.compute_rigid_body_mass(obj_id)
                self.rigid_body_is_dynamic[obj_id] = 1
            self.present_object.append(obj_id)

        for _ in self.rigid_blocks:
            raise NotImplementedError

    def _add_uniform(self, obj_id, points, velocity, density, material, is_dynamic, color):
        points = np.array(points, dtype=np.float32)
        n = points.shape[0]
        self.add_particles(obj_id, n, points,
                           np.tile(np.asarray(velocity, dtype=np.float32), (n, 1)),
                           np.full(n, density, dtype=np.float32), np.zeros(n, dtype=np.float32),
                           np.full(n, material, dtype=np.int32), np.full(n, is_dynamic, dtype=np.int32),
                           np.tile(np.asarray(color, dtype=np.int32), (n, 1)))

    def compute_rigid_body_mass(self, object_id: int) -> float:
        return self._engine.compute_rigid_body_mass(object_id)

    def compute_rigid_body_center_of_mass(self, object_id: int):
        """base_container.py:392-401 (unused by upstream's own solvers): density-weighted mean position of the
        dynamic particles of one object, computed on the host from the fields."""
        n = self.particle_num[None]
        sel = (self.particle_object_ids.to_numpy(n) == object_id) & (self.particle_is_dynamic.to_numpy(n) != 0)
        w = self.particle_densities.to_numpy(n)[sel].astype(np.float32) * np.float32(self.V0)
        return (self.particle_positions.to_numpy(n)[sel] * w[:, None]).sum(0) / w.sum()

    def copy_to_numpy(self, np_arr, src_arr):
        """base_container.py:562-565: the live rows of a field into a caller-owned array."""
        n = self.particle_num[None]
        np_arr[:n] = src_arr.to_numpy(n)

    # upstream's neighbourhood search is three calls (base_container.py:544-547); here the whole sort is one library
    # call, issued by reorder_particles() so that the upstream sequence keeps working unchanged
    def init_grid(self):
        pass

    def reorder_particles(self):
        self._engine.prepare_neighborhood_search()

    # host-side equivalents of upstream's device functions (@ti.func, base_container.py:467-493), for Python-side
    # tasks over neighbor_lists()
    def pos_to_index(self, pos):
        return (np.asarray(pos, dtype=np.float32) / np.float32(self.grid_size)).astype(np.int32)

    def flatten_grid_index(self, grid_index):
        g = np.asarray(grid_index, dtype=np.int64)
        return (g[..., 0] * int(self.grid_num[1]) + g[..., 1]) * int(self.grid_num[2]) + g[..., 2]

    def get_flatten_grid_index(self, pos):
        return self.flatten_grid_index(self.pos_to_index(pos))

    def is_static_rigid_body(self, p):
        return bool(self.particle_materials[p] == self.material_rigid and not self.particle_is_dynamic[p])

    def is_dynamic_rigid_body(self, p):
        return bool(self.particle_materials[p] == self.material_rigid and self.particle_is_dynamic[p])

    def add_particles(self, object_id, new_particles_num, new_particles_positions, new_particles_velocity,
                      new_particle_density, new_particle_pressure, new_particles_material,
                      new_particles_is_dynamic, new_particles_color):
        """Append particles (base_container.py:417-464); host arrays are copied to the device."""
        positions = np.asarray(new_particles_positions, dtype=np.float32).reshape(-1, self.dim)
        assert new_particles_num == positions.shape[0]
        if self.slab is None:
            self._engine.add_particles(object_id, positions, new_particles_velocity, new_particle_density,
                                       new_particle_pressure, new_particles_material, new_particles_is_dynamic,
                                       new_particles_color)
            self._next_uid += new_particles_num
            return
        keep = self._slab_range_now().owned(positions)
        pick = lambda a, w: np.asarray(a).reshape(new_particles_num, *([w] if w > 1 else []))[keep]
        self._add_slab_rows(object_id, new_particles_num, np.flatnonzero(keep), positions[keep],
                            pick(new_particles_velocity, self.dim), pick(new_particle_density, 1),
                            pick(new_particle_pressure, 1), pick(new_particles_material, 1),
                            pick(new_particles_is_dynamic, 1), pick(new_particles_color, 3))

    def _slab_range_now(self):
        """My layers as they are now: the library may have moved the boundaries towards the busier rank since the
        scene was cut."""
        info = self._engine.slab_info()
        self.slab.ranges[self.slab.rank] = (info.z_lo, info.z_hi)
        return self.slab

    def _slab_z_indices(self, z_axis):
        """Indices of the lattice z coordinates that fall into my layers."""
        return np.flatnonzero(self._slab_range_now().owned_z(z_axis.astype(np.float32)))

    def _add_slab_rows(self, object_id, total_num, rows, positions, velocity, density, pressure, material, is_dynamic,
                       color):
        """Z-slab insertion of an object with `total_num` particles of which this rank keeps `rows` (row numbers within
        the object); uids stay the global insertion indices."""
        uids = (self._next_uid + np.asarray(rows, dtype=np.int64)).astype(np.int32)
        self._next_uid += total_num
        # particle_num of the whole domain as of now (the DFSPH error averages over it, DFSPH.py:211,294): every rank
        # sees every insertion, whoever keeps the particles
        self._engine.slab_set_global_particle_num(self._next_uid)
        if not len(uids):
            return
        n_before = self.particle_num[None]
        self._engine.add_particles(object_id, positions, velocity, density, pressure, material, is_dynamic, color)
        all_uids = np.concatenate([self._engine.get_field(F.UID, n_before), uids])
        self._engine.set_field(F.UID, all_uids)

    def _add_lattice(self, object_id, positions, material, is_dynamic, color, density, pressure, velocity, slab_rows=None):
        """Insert lattice points with uniform attributes.  slab_rows = (total count of the object, row numbers of
        `positions` within it): a Z-slab rank passes the rows of its own layers only."""
        n = positions.shape[0]
        if velocity is None:
            velocity_arr = np.zeros_like(positions, dtype=np.float32)
        else:
            velocity_arr = np.tile(np.asarray(velocity, dtype=np.float32), (n, 1))
        fields = (velocity_arr,
                  np.full(n, density if density is not None else 1000.0, dtype=np.float32),
                  np.full(n, pressure if pressure is not None else 0.0, dtype=np.float32),
                  np.full(n, material, dtype=np.int32), np.full(n, int(is_dynamic), dtype=np.int32),
                  np.tile(np.asarray(color, dtype=np.int32), (n, 1)))
        if slab_rows is None:
            self.add_particles(object_id, n, positions, *fields)
            return n
        total, rows = slab_rows
        self._add_slab_rows(object_id, total, rows, positions, *fields)
        return total

    def add_cube(self, object_id, lower_corner, cube_size, material, is_dynamic, color=(0, 0, 0), density=None,
                 pressure=None, velocity=None, space=None):
        """Particles spaced by `space` filling a box (base_container.py:753-798)."""
        if space is None:
            space = self.particle_diameter
        if self.slab is None:
            positions = _lattice(lower_corner, cube_size, space, self.dim)
            n = self._add_lattice(object_id, positions, material, is_dynamic, color, density, pressure, velocity)
        else:
            axes = _lattice_axes(lower_corner, cube_size, space, self.dim)
            positions, rows = _lattice_rows(axes, self._slab_z_indices(axes[2]))
            n = self._add_lattice(object_id, positions, material, is_dynamic, color, density, pressure, velocity,
                                  slab_rows=(len(axes[0]) * len(axes[1]) * len(axes[2]), rows))
        if material == self.material_fluid:
            self.fluid_particle_num[None] += n

    def _box_shell(self, lower_corner, cube_size, thickness, space):
        positions = _lattice(lower_corner, cube_size, space, self.dim)
        mask = np.zeros(positions.shape[0], dtype=bool)
        for i in range(self.dim):  # keep the shell, drop the interior (:830-835)
            mask |= (positions[:, i] <= lower_corner[i] + thickness) | \
                    (positions[:, i] >= lower_corner[i] + cube_size[i] - thickness)
        return positions[mask]

    def add_box(self, object_id, lower_corner, cube_size, thickness, material, is_dynamic, color=(0, 0, 0),
                density=None, pressure=None, velocity=None, space=None):
        """Hollow box of boundary particles (base_container.py:800-849)."""
        if space is None:
            space = self.particle_diameter
        if self.slab is None:
            positions = self._box_shell(lower_corner, cube_size, thickness, space)
            self._add_lattice(object_id, positions, material, is_dynamic, color, density, pressure, velocity)
            return
        axes = _lattice_axes(lower_corner, cube_size, space, self.dim)
        masks = _shell_masks(axes, lower_corner, cube_size, thickness)
        positions, rows = _shell_rows(axes, masks, self._slab_z_indices(axes[2]))
        self._add_lattice(object_id, positions, material, is_dynamic, color, density, pressure, velocity,
                          slab_rows=(int(_shell_z_counts(masks).sum()), rows))

    def compute_cube_particle_num(self, start, end, space=None):
        if space is None:
            space = self.particle_diameter
        return reduce(lambda a, b: a * b, [len(np.arange(start[i], end[i], space)) for i in range(self.dim)])

    def compute_box_particle_num(self, lower_corner, cube_size, thickness, space=None):
        if space is None:
            space = self.particle_diameter
        if self.dim != 3:
            return self._box_shell(lower_corner, cube_size, thickness, space).shape[0]
        axes = _lattice_axes(lower_corner, cube_size, space, self.dim)   # counted per axis: no lattice is built
        return int(_shell_z_counts(_shell_masks(axes, lower_corner, cube_size, thickness)).sum())

    # ------------------------------------------------------------------ mesh bodies
    def load_rigid_body(self, rigid_body, pitch=None):
        from ..mesh import voxelize_rigid_body
        return voxelize_rigid_body(rigid_body, pitch if pitch is not None else self.particle_diameter)

    def load_fluid_body(self, fluid_body, pitch=None):
        from ..mesh import voxelize_fluid_body
        return voxelize_fluid_body(fluid_body, pitch if pitch is not None else self.particle_diameter, self.dim)

    # ------------------------------------------------------------------ neighbourhood search
    def prepare_neighborhood_search(self):
        """init_grid + prefix sum + reorder (base_container.py:544-547) as CUDA kernels."""
        self._engine.prepare_neighborhood_search()

    @property
    def grid_num_particles(self):
        """Inclusive scan of per-cell counts in the reference's z-fastest cell order (:132,546)."""
        return _GridCountsView(self._engine)

    def neighbor_lists(self):
        """CSR (offsets, indices) of N(i) for every particle in current order."""
        return self._engine.get_neighbors()

    def for_all_neighbors(self, p_i, task, ret=None):
        """Host-side mirror of the device callback idiom (base_container.py:549-560).

        On the device each upstream `*_task` is a C++ functor compiled into the sweep kernels; this
        method is for custom Python tasks and tests: it calls `task(p_i, p_j, ret)` for every
        neighbour and returns the accumulated `ret` (Python scalars cannot be passed by reference,
        so a task may return the new value)."""
        offsets, indices = self._engine.get_neighbors()
        for p_j in indices[offsets[p_i]:offsets[p_i + 1]]:
            out = task(p_i, int(p_j), ret)
            if out is not None:
                ret = out
        return ret

    # ------------------------------------------------------------------ visualisation / export
    def copy_to_vis_buffer(self, invisible_objects=[], dim=3):
        assert self.GGUI
        self.flush_vis_buffer()
        ids = self.particle_object_ids.to_numpy()
        pos = self.particle_positions.to_numpy()
        col = self.particle_colors.to_numpy()
        for obj_id in self.object_collection:
            if self.object_visibility[obj_id] == 1:
                m = ids == obj_id
                self.x_vis_buffer[m] = pos[m]
                self.color_vis_buffer[m] = col[m] / 255.0

    def flush_vis_buffer(self):
        self.x_vis_buffer.fill(0.0)
        self.color_vis_buffer.fill(0.0)

    def dump(self, obj_id):
        """Positions and velocities of one object as host arrays (base_container.py:599-609)."""
        mask = (self.particle_object_ids.to_numpy() == obj_id).nonzero()
        return {
            "position": self.particle_positions.to_numpy()[mask],
            "velocity": self.particle_velocities.to_numpy()[mask],
        }
