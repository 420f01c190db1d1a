"""ctypes binding of the C ABI declared in include/sph_b200.h.

The product always binds the CUDA library (sph_project_b200/csrc/libsph_b200.so) and fails
loudly when it is missing: there is no CPU fallback.  `Engine` takes an already-loaded
`ctypes.CDLL` so that tests can drive the very same Python surface over any library exporting
the ABI (the parity tests do that with the CPU oracle, which lives outside this package).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

ABI_VERSION = 5
MAX_OBJECTS = 20

METHOD_WCSPH, METHOD_PCISPH, METHOD_DFSPH = 0, 1, 2
VISC_STANDARD, VISC_IMPLICIT = 0, 1
MATERIAL_FLUID, MATERIAL_RIGID = 1, 2
FLAG_SLAB = 1
SLAB_PEER_BLOB_BYTES = 512


class SphParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dim", C.c_int32), ("method", C.c_int32), ("visc_method", C.c_int32),
        ("max_particles", C.c_int32), ("grid_num", C.c_int32 * 3),
        ("dx", C.c_double), ("dh", C.c_double), ("V0", C.c_double), ("density0", C.c_double), ("dt", C.c_double),
        ("gravity", C.c_double * 3), ("g_upper", C.c_double), ("viscosity", C.c_double), ("viscosity_b", C.c_double),
        ("surface_tension", C.c_double), ("domain_size", C.c_double * 3), ("padding", C.c_double),
        ("device", C.c_int32), ("flags", C.c_int32),
    ]


class SphStepStats(C.Structure):
    _fields_ = [
        ("steps", C.c_int32), ("dfsph_iterations", C.c_int32), ("dfsph_iterations_v", C.c_int32),
        ("pcisph_iterations", C.c_int32), ("cg_iterations", C.c_int32),
        ("dfsph_density_error", C.c_float), ("dfsph_divergence_error", C.c_float),
        ("pcisph_density_error", C.c_float), ("cg_error", C.c_float),
        ("total_dfsph_iterations", C.c_int64), ("total_dfsph_iterations_v", C.c_int64),
        ("total_pcisph_iterations", C.c_int64), ("total_cg_iterations", C.c_int64),
        ("kernel_launches", C.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class SphKernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 56), ("launches", C.c_int64), ("total_ms", C.c_double)]


class SphSlabInfo(C.Structure):
    _fields_ = [("z_lo", C.c_int32), ("z_hi", C.c_int32), ("n_owned", C.c_int32), ("n_ghost", C.c_int32),
                ("n_send_lo", C.c_int32), ("n_send_hi", C.c_int32), ("own_begin", C.c_int32), ("own_end", C.c_int32),
                ("halo_bytes", C.c_int64), ("halo_calls", C.c_int64)]


class F:
    """SphField ids."""
    OBJECT_ID, POSITION, VELOCITY, ACCELERATION, REST_VOLUME, MASS, DENSITY, PRESSURE = range(8)
    MATERIAL, COLOR, IS_DYNAMIC, ORIGINAL_POSITION, GRID_ID, UID, CELL = range(8, 15)
    DFSPH_ALPHA, DFSPH_KAPPA, DFSPH_KAPPA_V, DENSITY_STAR, DENSITY_DERIVATIVE = range(20, 25)
    PRESSURE_ACCELERATION, PREDICTED_VELOCITY, PREDICTED_POSITION = 30, 31, 32
    CG_P, ORIGINAL_VELOCITY, CG_AP, CG_X, CG_B, CG_R, CG_DIAG_INV = range(40, 47)
    NEIGHBOR_COUNT = 50


# field id -> (components, numpy dtype)
FIELD_LAYOUT = {
    F.OBJECT_ID: (1, np.int32), F.POSITION: (3, np.float32), F.VELOCITY: (3, np.float32),
    F.ACCELERATION: (3, np.float32), F.REST_VOLUME: (1, np.float32), F.MASS: (1, np.float32),
    F.DENSITY: (1, np.float32), F.PRESSURE: (1, np.float32), F.MATERIAL: (1, np.int32), F.COLOR: (3, np.int32),
    F.IS_DYNAMIC: (1, np.int32), F.ORIGINAL_POSITION: (3, np.float32), F.GRID_ID: (1, np.int32),
    F.UID: (1, np.int32), F.CELL: (3, np.int32),
    F.DFSPH_ALPHA: (1, np.float32), F.DFSPH_KAPPA: (1, np.float32), F.DFSPH_KAPPA_V: (1, np.float32),
    F.DENSITY_STAR: (1, np.float32), F.DENSITY_DERIVATIVE: (1, np.float32),
    F.PRESSURE_ACCELERATION: (3, np.float32), F.PREDICTED_VELOCITY: (3, np.float32),
    F.PREDICTED_POSITION: (3, np.float32),
    F.CG_P: (3, np.float32), F.ORIGINAL_VELOCITY: (3, np.float32), F.CG_AP: (3, np.float32),
    F.CG_X: (3, np.float32), F.CG_B: (3, np.float32), F.CG_R: (3, np.float32), F.CG_DIAG_INV: (9, np.float32),
    F.NEIGHBOR_COUNT: (1, np.int32),
}


class S:
    """SphScalar ids."""
    DT, PARTICLE_NUM, FLUID_PARTICLE_NUM, PCISPH_K, DENSITY_ERROR, CG_ALPHA, CG_BETA, CG_ERROR = range(8)
    G_UPPER, VISCOSITY, VISCOSITY_B, NUM_CELLS, MAX_PARTICLES = range(8, 13)
    ACTIVE_BRICKS, MAX_WINDOW_SLOTS, WINDOW_OVERFLOWS = 13, 14, 15   # diagnostics of the brick-tile sweeps


class T:
    """SphTask ids (one per upstream @ti.kernel)."""
    COMPUTE_RIGID_PARTICLE_VOLUME = 0
    COMPUTE_PRESSURE_ACCELERATION = 1
    COMPUTE_GRAVITY_ACCELERATION = 2
    COMPUTE_SURFACE_TENSION_ACCELERATION = 3
    COMPUTE_VISCOSITY_ACCELERATION_STANDARD = 4
    COMPUTE_DENSITY = 5
    ENFORCE_DOMAIN_BOUNDARY_3D = 6
    RENEW_RIGID_PARTICLE_STATE = 7
    UPDATE_FLUID_VELOCITY = 8
    UPDATE_FLUID_POSITION = 9
    PREPARE_EMITTER = 10
    INIT_OBJECT_ID = 11
    INIT_ACCELERATION = 12
    INIT_RIGID_BODY_FORCE_AND_TORQUE = 13
    CG_PREPARE1 = 20
    CG_PREPARE2 = 21
    CG_COMPUTE_AP = 22
    CG_COMPUTE_ALPHA = 23
    CG_UPDATE_X = 24
    CG_UPDATE_R_AND_BETA = 25
    CG_UPDATE_P = 26
    CG_PREPARE_GUESS = 27
    VISCOSITY_UPDATE_VELOCITY = 28
    COPY_BACK_ORIGINAL_VELOCITY = 29
    WCSPH_COMPUTE_PRESSURE = 40
    DFSPH_COMPUTE_ALPHA = 50
    DFSPH_COMPUTE_DENSITY_DERIVATIVE = 51
    DFSPH_COMPUTE_DENSITY_STAR = 52
    DFSPH_COMPUTE_KAPPA_V = 53
    DFSPH_CORRECT_DIVERGENCE_STEP = 54
    DFSPH_COMPUTE_DENSITY_DERIVATIVE_ERROR = 55
    DFSPH_COMPUTE_KAPPA = 56
    DFSPH_CORRECT_DENSITY_ERROR_STEP = 57
    DFSPH_COMPUTE_DENSITY_ERROR = 58
    PCISPH_COMPUTE_PREDICTED_VELOCITY = 70
    PCISPH_COMPUTE_PREDICTED_POSITION = 71
    PCISPH_COMPUTE_DENSITY_STAR = 72
    PCISPH_UPDATE_PRESSURE = 73
    PCISPH_COMPUTE_TEMP_PRESSURE_ACCELERATION = 74
    PCISPH_COMPUTE_K = 75
    PCISPH_INIT_STEP = 76


ERROR_NAMES = {0: "SPH_OK", -1: "SPH_E_INVALID", -2: "SPH_E_CAPACITY", -3: "SPH_E_CUDA", -4: "SPH_E_STATE",
               -5: "SPH_E_UNSUPPORTED", -6: "SPH_E_NOMEM"}

_H = C.c_void_p
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)

# every symbol include/sph_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "sph_create": (C.c_int, [C.POINTER(SphParams), C.POINTER(_H)]),
    "sph_destroy": (C.c_int, [_H]),
    "sph_last_error": (C.c_char_p, [_H]),
    "sph_backend_name": (C.c_char_p, []),
    "sph_abi_version": (C.c_int, []),
    "sph_add_particles": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "sph_get_field": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_size_t]),
    "sph_set_field": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_size_t]),
    "sph_fill_field": (C.c_int, [_H, C.c_int32, C.c_double]),
    "sph_get_scalar": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_double)]),
    "sph_set_scalar": (C.c_int, [_H, C.c_int32, C.c_double]),
    "sph_field_ptr": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_void_p), _i32p, _i32p]),
    "sph_set_object": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "sph_set_rigid_state": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sph_get_rigid_wrench": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "sph_zero_rigid_wrench": (C.c_int, [_H]),
    "sph_compute_rigid_body_mass": (C.c_int, [_H, C.c_int32, _f32p]),
    "sph_prepare_neighborhood_search": (C.c_int, [_H]),
    "sph_get_neighbors": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_size_t]),
    "sph_get_grid_num_particles": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "sph_run_task": (C.c_int, [_H, C.c_int32, C.c_int32, _f32p]),
    "sph_step": (C.c_int, [_H, C.c_int32, C.POINTER(SphStepStats)]),
    "sph_dfsph_correct_density_error": (C.c_int, [_H, _i32p, _f32p]),
    "sph_dfsph_correct_divergence_error": (C.c_int, [_H, _i32p, _f32p]),
    "sph_pcisph_refine": (C.c_int, [_H, _i32p, _f32p]),
    "sph_implicit_viscosity_solve": (C.c_int, [_H, _i32p, _f32p]),
    "sph_synchronize": (C.c_int, [_H]),
    "sph_set_stream": (C.c_int, [_H, C.c_void_p]),
    "sph_profile_enable": (C.c_int, [_H, C.c_int32]),
    "sph_profile_read": (C.c_int, [_H, C.POINTER(SphKernelStat), C.c_int32, _i32p]),
    "sph_slab_unique_id": (C.c_int, [C.c_void_p]),
    "sph_slab_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int64]),
    "sph_slab_set_global_particle_num": (C.c_int, [_H, C.c_int64]),
    "sph_slab_info": (C.c_int, [_H, C.POINTER(SphSlabInfo)]),
    "sph_slab_peer_export": (C.c_int, [_H, C.c_void_p]),
    "sph_slab_peer_import": (C.c_int, [_H, C.c_int32, C.c_void_p]),
}


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach the ABI prototypes to a loaded library; raises AttributeError on a missing symbol."""
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sph_abi_version() != ABI_VERSION:
        raise RuntimeError(f"ABI mismatch: library has {lib.sph_abi_version()}, binding expects {ABI_VERSION}")
    return lib


CUDA_LIBRARY_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libsph_b200.so")
_cuda_lib: Optional[C.CDLL] = None


def load_cuda_library() -> C.CDLL:
    """Load the sm_100a CUDA library.  No fallback: a missing build is an error."""
    global _cuda_lib
    if _cuda_lib is None:
        # SPH_B200_LIBRARY selects another build of the SAME CUDA library (a tuning variant from `make variants`);
        # it is still sm_100a-only and must report the CUDA backend
        path = os.environ.get("SPH_B200_LIBRARY") or CUDA_LIBRARY_PATH
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C sph_project_b200/csrc`).  sph_project_b200 has no CPU fallback.")
        lib = bind(C.CDLL(path))
        if lib.sph_backend_name().decode() != "cuda-sm100a":
            raise RuntimeError(f"{path} is not the sm_100a CUDA library (backend {lib.sph_backend_name().decode()!r})")
        _cuda_lib = lib
    return _cuda_lib


class SphError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One SphHandle.  Thin, typed access to the C ABI; all numpy in/out is host memory."""

    def __init__(self, params: SphParams, lib: Optional[C.CDLL] = None):
        self.lib = lib if lib is not None else load_cuda_library()
        self.params = params
        self._h = _H()
        rc = self.lib.sph_create(C.byref(params), C.byref(self._h))
        if rc != 0:
            raise SphError(rc, "sph_create failed (is a B200 visible? are the parameters valid?)")
        self.backend = self.lib.sph_backend_name().decode()

    # -- plumbing --
    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.sph_last_error(self._h)
            raise SphError(rc, msg.decode() if msg else "")

    def close(self):
        if self._h:
            self.lib.sph_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scalars --
    def get_scalar(self, sid: int) -> float:
        out = C.c_double()
        self._check(self.lib.sph_get_scalar(self._h, sid, C.byref(out)))
        return out.value

    def set_scalar(self, sid: int, value: float):
        self._check(self.lib.sph_set_scalar(self._h, sid, float(value)))

    @property
    def particle_num(self) -> int:
        return int(self.get_scalar(S.PARTICLE_NUM))

    # -- fields --
    def get_field(self, fid: int, n: Optional[int] = None) -> np.ndarray:
        comps, dtype = FIELD_LAYOUT[fid]
        if n is None:
            n = self.particle_num
        shape = (n,) if comps == 1 else (n, comps)
        out = np.empty(shape, dtype=dtype)
        if n:
            self._check(self.lib.sph_get_field(self._h, fid, _ptr(out), out.nbytes))
        return out

    def get_field_into(self, fid: int, out: np.ndarray):
        self._check(self.lib.sph_get_field(self._h, fid, _ptr(out), out.nbytes))

    def set_field(self, fid: int, values: np.ndarray):
        comps, dtype = FIELD_LAYOUT[fid]
        a = np.ascontiguousarray(values, dtype=dtype)
        if a.size:
            self._check(self.lib.sph_set_field(self._h, fid, _ptr(a), a.nbytes))

    def fill_field(self, fid: int, value: float):
        self._check(self.lib.sph_fill_field(self._h, fid, float(value)))

    def field_ptr(self, fid: int) -> Tuple[int, int, int]:
        p, stride, comps = C.c_void_p(), C.c_int32(), C.c_int32()
        self._check(self.lib.sph_field_ptr(self._h, fid, C.byref(p), C.byref(stride), C.byref(comps)))
        return int(p.value or 0), stride.value, comps.value

    # -- particles / objects --
    def add_particles(self, object_id, x, v, density, pressure, material, is_dynamic, color):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
        n = x.shape[0]
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(n, 3)
        density = np.ascontiguousarray(density, dtype=np.float32).reshape(n)
        pressure = np.ascontiguousarray(pressure, dtype=np.float32).reshape(n)
        material = np.ascontiguousarray(material, dtype=np.int32).reshape(n)
        is_dynamic = np.ascontiguousarray(is_dynamic, dtype=np.int32).reshape(n)
        color = np.ascontiguousarray(color, dtype=np.int32).reshape(n, 3)
        self._check(self.lib.sph_add_particles(self._h, int(object_id), n, _ptr(x), _ptr(v), _ptr(density),
                                               _ptr(pressure), _ptr(material), _ptr(is_dynamic), _ptr(color)))

    def set_object(self, object_id: int, material: int, is_dynamic: int):
        self._check(self.lib.sph_set_object(self._h, int(object_id), int(material), int(is_dynamic)))

    def set_rigid_state(self, object_id, com0=None, com=None, rotation=None, velocity=None, angular_velocity=None):
        def f(a, n):
            return None if a is None else np.ascontiguousarray(a, dtype=np.float32).reshape(n)
        args = [f(com0, 3), f(com, 3), f(rotation, 9), f(velocity, 3), f(angular_velocity, 3)]
        self._check(self.lib.sph_set_rigid_state(self._h, int(object_id), *[_ptr(a) for a in args]))

    def get_rigid_wrench(self):
        force = np.zeros((MAX_OBJECTS, 3), dtype=np.float32)
        torque = np.zeros((MAX_OBJECTS, 3), dtype=np.float32)
        self._check(self.lib.sph_get_rigid_wrench(self._h, _ptr(force), _ptr(torque)))
        return force, torque

    def zero_rigid_wrench(self):
        self._check(self.lib.sph_zero_rigid_wrench(self._h))

    def compute_rigid_body_mass(self, object_id: int) -> float:
        out = C.c_float()
        self._check(self.lib.sph_compute_rigid_body_mass(self._h, int(object_id), C.byref(out)))
        return out.value

    # -- neighbourhood --
    def prepare_neighborhood_search(self):
        self._check(self.lib.sph_prepare_neighborhood_search(self._h))

    def get_neighbors(self):
        n = self.particle_num
        offsets = np.zeros(n + 1, dtype=np.int32)
        self._check(self.lib.sph_get_neighbors(self._h, _ptr(offsets), None, 0))
        indices = np.zeros(max(int(offsets[n]), 1), dtype=np.int32)
        self._check(self.lib.sph_get_neighbors(self._h, _ptr(offsets), _ptr(indices), indices.size))
        return offsets, indices[: offsets[n]]

    def get_grid_num_particles(self) -> np.ndarray:
        out = np.zeros(int(self.get_scalar(S.NUM_CELLS)), dtype=np.int32)
        self._check(self.lib.sph_get_grid_num_particles(self._h, _ptr(out), out.size))
        return out

    # -- kernels --
    def run_task(self, task: int, iarg: int = 0) -> float:
        out = C.c_float(0.0)
        self._check(self.lib.sph_run_task(self._h, int(task), int(iarg), C.byref(out)))
        return out.value

    def step(self, n_steps: int = 1) -> SphStepStats:
        st = SphStepStats()
        self._check(self.lib.sph_step(self._h, int(n_steps), C.byref(st)))
        return st

    def _loop(self, fn):
        it, err = C.c_int32(), C.c_float()
        self._check(fn(self._h, C.byref(it), C.byref(err)))
        return it.value, err.value

    def dfsph_correct_density_error(self):
        return self._loop(self.lib.sph_dfsph_correct_density_error)

    def dfsph_correct_divergence_error(self):
        return self._loop(self.lib.sph_dfsph_correct_divergence_error)

    def pcisph_refine(self):
        return self._loop(self.lib.sph_pcisph_refine)

    def implicit_viscosity_solve(self):
        return self._loop(self.lib.sph_implicit_viscosity_solve)

    def synchronize(self):
        self._check(self.lib.sph_synchronize(self._h))

    def set_stream(self, cuda_stream: int):
        """Run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream); 0 restores."""
        self._check(self.lib.sph_set_stream(self._h, C.c_void_p(cuda_stream or None)))

    def profile_enable(self, on: bool = True):
        self._check(self.lib.sph_profile_enable(self._h, int(on)))

    def profile_read(self):
        """{kernel name: (launches, total device ms)} since the last read."""
        rows = (SphKernelStat * 128)()
        count = C.c_int32()
        self._check(self.lib.sph_profile_read(self._h, rows, 128, C.byref(count)))
        return {rows[i].name.decode(): (rows[i].launches, rows[i].total_ms) for i in range(count.value)}

    # -- slabs --
    def slab_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.sph_slab_unique_id(buf)
        if rc != 0:
            raise SphError(rc, "sph_slab_unique_id (is NCCL loaded in this process?)")
        return buf.raw

    def slab_init(self, rank: int, world: int, unique_id: bytes, z_lo: int, z_hi: int, global_particle_num: int):
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.sph_slab_init(self._h, int(rank), int(world), buf, int(z_lo), int(z_hi), int(global_particle_num)))

    def slab_set_global_particle_num(self, n: int):
        self._check(self.lib.sph_slab_set_global_particle_num(self._h, int(n)))

    def slab_peer_export(self) -> bytes:
        buf = C.create_string_buffer(SLAB_PEER_BLOB_BYTES)
        self._check(self.lib.sph_slab_peer_export(self._h, buf))
        return buf.raw

    def slab_peer_import(self, peer_rank: int, blob: bytes):
        assert len(blob) == SLAB_PEER_BLOB_BYTES
        self._check(self.lib.sph_slab_peer_import(self._h, int(peer_rank), C.create_string_buffer(blob, SLAB_PEER_BLOB_BYTES)))

    def slab_info(self) -> SphSlabInfo:
        info = SphSlabInfo()
        self._check(self.lib.sph_slab_info(self._h, C.byref(info)))
        return info
