/*
 * sph_b200.h — C ABI of the B200-native SPH inner loop.
 *
 * The reference (jason-huang03/SPH_Project) has no FFI of its own: its boundary is the
 * Python class surface of SPH/containers + SPH/fluid_solvers on top of Taichi fields and
 * @ti.kernel launches.  This header is the thin C ABI that replaces the Taichi runtime
 * underneath that surface.  Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout).
 *
 * Two shared libraries export exactly these symbols:
 *   sph_project_b200/csrc/libsph_b200.so   hand-written sm_100a CUDA (the product)
 *   oracle/_build/libsph_oracle.so         CPU restatement of the reference (test infrastructure only)
 *
 * Conventions
 *   - every function returns 0 on success and a negative SPH_E_* code on failure; nothing
 *     throws or aborts across the ABI; sph_last_error() returns a message for the handle.
 *   - the caller owns host buffers; the library owns device buffers.  Host buffers passed to
 *     get/set are plain pointers + byte sizes, copied synchronously.
 *   - a handle is not thread-safe: one host thread per handle, one CUDA stream per handle.
 *   - particle arrays are in the library's CURRENT (cell-sorted) order; SPH_F_UID carries the
 *     insertion index of every particle so callers can undo the permutation.
 *   - 3-D only (the reference's 2-D branches are unreachable: rigid_solver/bullet_solver.py:19).
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_ABI_VERSION 5
#define SPH_MAX_OBJECTS 20 /* base_container.py:52 max_num_object */

/* error codes */
#define SPH_OK 0
#define SPH_E_INVALID (-1)   /* bad argument / unknown id */
#define SPH_E_CAPACITY (-2)  /* particle_max_num exceeded (base_container.py:116) */
#define SPH_E_CUDA (-3)      /* CUDA runtime error (message in sph_last_error) */
#define SPH_E_STATE (-4)     /* call not valid in the current state */
#define SPH_E_UNSUPPORTED (-5)
#define SPH_E_NOMEM (-6)

/* simulationMethod (run_simulation.py:45-63) */
#define SPH_METHOD_WCSPH 0
#define SPH_METHOD_PCISPH 1
#define SPH_METHOD_DFSPH 2
/* viscosityMethod (base_solver.py:195-200) */
#define SPH_VISC_STANDARD 0
#define SPH_VISC_IMPLICIT 1
/* particle materials (base_container.py:29-30) */
#define SPH_MATERIAL_FLUID 1
#define SPH_MATERIAL_RIGID 2

/* Scene constants; mirrors what BaseContainer.__init__ (base_container.py:10-66) and
 * BaseSolver.__init__ (base_solver.py:9-54) derive from the scene JSON. Python-side constants
 * are f64 there and meet f32 operands inside kernels, hence double here. */
typedef struct SphParams {
    int32_t abi_version;     /* = SPH_ABI_VERSION */
    int32_t dim;             /* must be 3 */
    int32_t method;          /* SPH_METHOD_* */
    int32_t visc_method;     /* SPH_VISC_* */
    int32_t max_particles;   /* particle_max_num */
    int32_t grid_num[3];     /* ceil(domain_size / dh), base_container.py:56 */
    double dx;               /* particleRadius */
    double dh;               /* support radius = grid cell size = padding */
    double V0;               /* 0.8 * diameter^3, base_container.py:49 */
    double density0;         /* base_solver.py:31 */
    double dt;               /* timeStepSize */
    double gravity[3];       /* gravitation */
    double g_upper;          /* gravitationUpper or 1e4, base_solver.py:21-23 */
    double viscosity;        /* base_solver.py:26 */
    double viscosity_b;      /* base_solver.py:27-29 */
    double surface_tension;  /* 0.01, base_solver.py:32 */
    double domain_size[3];   /* domainEnd - domainStart (domainStart must be 0) */
    double padding;          /* = dh, base_container.py:58 */
    int32_t device;          /* CUDA device ordinal (ignored by the oracle) */
    int32_t flags;           /* SPH_FLAG_* */
} SphParams;

#define SPH_FLAG_SLAB 1      /* handle is one Z-slab of a sharded domain (multi-GPU) */

typedef struct SphHandle SphHandle;

/* Per-particle fields. Layout on the host side of get/set is dense row-major
 * ([n], [n,3] or [n,9]) of f32 or i32, n = particle_num. */
typedef enum SphField {
    SPH_F_OBJECT_ID = 0,       /* i32      particle_object_ids            base_container.py:138 */
    SPH_F_POSITION = 1,        /* f32[3]   particle_positions             :139 */
    SPH_F_VELOCITY = 2,        /* f32[3]   particle_velocities            :140 */
    SPH_F_ACCELERATION = 3,    /* f32[3]   particle_accelerations         :141 */
    SPH_F_REST_VOLUME = 4,     /* f32      particle_rest_volumes          :142 */
    SPH_F_MASS = 5,            /* f32      particle_masses                :143 */
    SPH_F_DENSITY = 6,         /* f32      particle_densities             :144 */
    SPH_F_PRESSURE = 7,        /* f32      particle_pressures             :145 */
    SPH_F_MATERIAL = 8,        /* i32      particle_materials             :146 */
    SPH_F_COLOR = 9,           /* i32[3]   particle_colors                :147 */
    SPH_F_IS_DYNAMIC = 10,     /* i32      particle_is_dynamic            :148 */
    SPH_F_ORIGINAL_POSITION = 11, /* f32[3] rigid_particle_original_positions :155 */
    SPH_F_GRID_ID = 12,        /* i32      grid_ids (library's own flatten) :183 */
    SPH_F_UID = 13,            /* i32      insertion index (not in the reference) */
    SPH_F_CELL = 14,           /* i32[3]   cell coordinate trunc(x/dh)  pos_to_index :467-469 (get only) */
    /* DFSPH (dfsph_container.py:13-17) */
    SPH_F_DFSPH_ALPHA = 20,
    SPH_F_DFSPH_KAPPA = 21,
    SPH_F_DFSPH_KAPPA_V = 22,
    SPH_F_DENSITY_STAR = 23,   /* shared with PCISPH (pcisph_container.py:19) */
    SPH_F_DENSITY_DERIVATIVE = 24,
    /* PCISPH (pcisph_container.py:16-18) */
    SPH_F_PRESSURE_ACCELERATION = 30, /* f32[3] */
    SPH_F_PREDICTED_VELOCITY = 31,    /* f32[3] */
    SPH_F_PREDICTED_POSITION = 32,    /* f32[3] */
    /* implicit viscosity CG scratch (base_solver.py:43-52) */
    SPH_F_CG_P = 40,
    SPH_F_ORIGINAL_VELOCITY = 41,
    SPH_F_CG_AP = 42,
    SPH_F_CG_X = 43,
    SPH_F_CG_B = 44,
    SPH_F_CG_R = 45,
    SPH_F_CG_DIAG_INV = 46,    /* f32[9] row-major 3x3 */
    SPH_F_NEIGHBOR_COUNT = 50  /* i32 |N(i)| for every particle (debug; get only) */
} SphField;

/* Scalars (0-d Taichi fields / Python attributes in the reference). */
typedef enum SphScalar {
    SPH_S_DT = 0,                 /* solver.dt[None]                 base_solver.py:34-36 */
    SPH_S_PARTICLE_NUM = 1,       /* container.particle_num[None]    base_container.py:50 */
    SPH_S_FLUID_PARTICLE_NUM = 2, /* container.fluid_particle_num[None] :125 */
    SPH_S_PCISPH_K = 3,           /* container.pcisph_k[None]        pcisph_container.py:15 */
    SPH_S_DENSITY_ERROR = 4,      /* container.density_error[None]   pcisph_container.py:14 */
    SPH_S_CG_ALPHA = 5,           /* base_solver.py:48 */
    SPH_S_CG_BETA = 6,            /* base_solver.py:49 */
    SPH_S_CG_ERROR = 7,           /* base_solver.py:51 */
    SPH_S_G_UPPER = 8,
    SPH_S_VISCOSITY = 9,
    SPH_S_VISCOSITY_B = 10,
    SPH_S_NUM_CELLS = 11,
    SPH_S_MAX_PARTICLES = 12,
    /* diagnostics of the brick-tile sweeps (no reference counterpart; the oracle reports 0) */
    SPH_S_ACTIVE_BRICKS = 13,     /* bricks that own fluid rows as of the last sort */
    SPH_S_MAX_WINDOW_SLOTS = 14,  /* largest brick window (particles) seen since the last sort */
    SPH_S_WINDOW_OVERFLOWS = 15   /* brick windows above the shared-memory budget since the last sort */
} SphScalar;

/* One id per upstream @ti.kernel on the hot path.  sph_run_task launches
 * exactly that kernel; `iarg` is its integer argument if it has one, `out` receives its
 * return value if it has one (else may be NULL). */
typedef enum SphTask {
    /* BaseSolver (SPH/fluid_solvers/base_solver.py) */
    SPH_T_COMPUTE_RIGID_PARTICLE_VOLUME = 0,   /* :105-123 */
    SPH_T_COMPUTE_PRESSURE_ACCELERATION = 1,   /* :135-187 */
    SPH_T_COMPUTE_GRAVITY_ACCELERATION = 2,    /* :202-207 */
    SPH_T_COMPUTE_SURFACE_TENSION_ACCELERATION = 3, /* :209-229 */
    SPH_T_COMPUTE_VISCOSITY_ACCELERATION_STANDARD = 4, /* :231-278 */
    SPH_T_COMPUTE_DENSITY = 5,                 /* :521-541 */
    SPH_T_ENFORCE_DOMAIN_BOUNDARY_3D = 6,      /* :574-605, iarg = particle_type */
    SPH_T_RENEW_RIGID_PARTICLE_STATE = 7,      /* :615-629 */
    SPH_T_UPDATE_FLUID_VELOCITY = 8,           /* :642-649 */
    SPH_T_UPDATE_FLUID_POSITION = 9,           /* :651-666 */
    SPH_T_PREPARE_EMITTER = 10,                /* :669-677 */
    SPH_T_INIT_OBJECT_ID = 11,                 /* :679-681 */
    SPH_T_INIT_ACCELERATION = 12,              /* :125-127 */
    SPH_T_INIT_RIGID_BODY_FORCE_AND_TORQUE = 13, /* :129-132 */
    /* implicit viscosity */
    SPH_T_CG_PREPARE1 = 20,                    /* prepare_conjugate_gradient_solver1 :281-315 */
    SPH_T_CG_PREPARE2 = 21,                    /* prepare_conjugate_gradient_solver2 :317-323 */
    SPH_T_CG_COMPUTE_AP = 22,                  /* :373-391 */
    SPH_T_CG_COMPUTE_ALPHA = 23,               /* :393-406 */
    SPH_T_CG_UPDATE_X = 24,                    /* :408-412 */
    SPH_T_CG_UPDATE_R_AND_BETA = 25,           /* :414-431, out = cg_error */
    SPH_T_CG_UPDATE_P = 26,                    /* :433-437 */
    SPH_T_CG_PREPARE_GUESS = 27,               /* :439-443 */
    SPH_T_VISCOSITY_UPDATE_VELOCITY = 28,      /* :463-467 */
    SPH_T_COPY_BACK_ORIGINAL_VELOCITY = 29,    /* :469-473 */
    /* WCSPH (SPH/fluid_solvers/WCSPH.py) */
    SPH_T_WCSPH_COMPUTE_PRESSURE = 40,         /* :16-24 */
    /* DFSPH (SPH/fluid_solvers/DFSPH.py) */
    SPH_T_DFSPH_COMPUTE_ALPHA = 50,            /* :22-62 */
    SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE = 51, /* :65-101 */
    SPH_T_DFSPH_COMPUTE_DENSITY_STAR = 52,     /* :104-126 */
    SPH_T_DFSPH_COMPUTE_KAPPA_V = 53,          /* :132-137 */
    SPH_T_DFSPH_CORRECT_DIVERGENCE_STEP = 54,  /* :161-202 */
    SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE_ERROR = 55, /* :205-211, out = error */
    SPH_T_DFSPH_COMPUTE_KAPPA = 56,            /* :217-223 */
    SPH_T_DFSPH_CORRECT_DENSITY_ERROR_STEP = 57, /* :245-283 */
    SPH_T_DFSPH_COMPUTE_DENSITY_ERROR = 58,    /* :285-294, out = error */
    /* PCISPH (SPH/fluid_solvers/PCISPH.py) */
    SPH_T_PCISPH_COMPUTE_PREDICTED_VELOCITY = 70, /* :18-22 */
    SPH_T_PCISPH_COMPUTE_PREDICTED_POSITION = 71, /* :25-29 */
    SPH_T_PCISPH_COMPUTE_DENSITY_STAR = 72,    /* :32-62 */
    SPH_T_PCISPH_UPDATE_PRESSURE = 73,         /* :65-71 */
    SPH_T_PCISPH_COMPUTE_TEMP_PRESSURE_ACCELERATION = 74, /* :74-107 */
    SPH_T_PCISPH_COMPUTE_K = 75,               /* compute_pcisph_k :128-151 */
    SPH_T_PCISPH_INIT_STEP = 76                /* :153-162 */
} SphTask;

/* Iteration counts and final errors of one solver step; what the reference prints at
 * DFSPH.py:159,243, PCISPH.py:125 and base_solver.py:461. */
typedef struct SphStepStats {
    int32_t steps;                 /* steps executed by this call */
    int32_t dfsph_iterations;      /* constant-density solve, last step */
    int32_t dfsph_iterations_v;    /* divergence-free solve, last step */
    int32_t pcisph_iterations;     /* last step */
    int32_t cg_iterations;         /* last step */
    float dfsph_density_error;     /* avg (rho_star/rho0 - 1), last step */
    float dfsph_divergence_error;  /* avg rho0 * D rho/Dt, last step */
    float pcisph_density_error;
    float cg_error;
    int64_t total_dfsph_iterations;   /* sums over all steps of this call */
    int64_t total_dfsph_iterations_v;
    int64_t total_pcisph_iterations;
    int64_t total_cg_iterations;
    int64_t kernel_launches;       /* device kernels launched by this call (0 for the oracle) */
} SphStepStats;

/* ---- lifetime ---------------------------------------------------------------------- */
/* Replaces BaseContainer.__init__ field allocation (base_container.py:129-190) and the
 * per-solver container add-ons ({wcsph,pcisph,dfsph}_container.py) + CG scratch
 * (base_solver.py:40-54). */
int sph_create(const SphParams* params, SphHandle** out);
int sph_destroy(SphHandle* h);
const char* sph_last_error(const SphHandle* h);
/* "cuda-sm100a" or "oracle-cpu" */
const char* sph_backend_name(void);
int sph_abi_version(void);

/* ---- particles --------------------------------------------------------------------- */
/* BaseContainer._add_particles / add_particle (base_container.py:403-464): appends n
 * particles at [particle_num, particle_num+n); rest_volume = V0, mass = V0*density,
 * original_position = x.  color is i32[n,3].  Does NOT touch fluid_particle_num
 * (the reference bumps that in add_cube / insert_object on the Python side). */
int sph_add_particles(SphHandle* h, int32_t object_id, int32_t n,
                      const float* x, const float* v, const float* density,
                      const float* pressure, const int32_t* material,
                      const int32_t* is_dynamic, const int32_t* color);

/* field[i] / field.to_numpy() / field.from_numpy() on the reference's Taichi fields. */
int sph_get_field(SphHandle* h, int32_t field, void* dst, size_t bytes);
int sph_set_field(SphHandle* h, int32_t field, const void* src, size_t bytes);
int sph_fill_field(SphHandle* h, int32_t field, double value);
int sph_get_scalar(SphHandle* h, int32_t scalar, double* out);
int sph_set_scalar(SphHandle* h, int32_t scalar, double value);
/* Raw device pointer + element stride (bytes) of a field for zero-copy wrapping
 * (torch / __cuda_array_interface__).  The oracle returns host pointers. */
int sph_field_ptr(SphHandle* h, int32_t field, void** ptr, int32_t* stride_bytes,
                  int32_t* components);

/* object tables: object_materials, rigid_body_is_dynamic (base_container.py:150,156) */
int sph_set_object(SphHandle* h, int32_t object_id, int32_t material, int32_t is_dynamic);
/* rigid_body_{original_centers_of_mass,centers_of_mass,rotations,velocities,
 * angular_velocities} written by bullet_solver.py:117-122,158-167 */
int sph_set_rigid_state(SphHandle* h, int32_t object_id, const float com0[3],
                        const float com[3], const float rotation[9],
                        const float velocity[3], const float angular_velocity[3]);
/* rigid_body_forces / rigid_body_torques read + zeroed by bullet_solver.py:149-156 */
int sph_get_rigid_wrench(SphHandle* h, float* force /*[20*3]*/, float* torque /*[20*3]*/);
int sph_zero_rigid_wrench(SphHandle* h);
/* BaseContainer.compute_rigid_body_mass (base_container.py:384-390) */
int sph_compute_rigid_body_mass(SphHandle* h, int32_t object_id, float* out);

/* ---- neighbourhood search ----------------------------------------------------------- */
/* BaseContainer.prepare_neighborhood_search (base_container.py:544-547):
 * init_grid + prefix sum + reorder_particles. */
int sph_prepare_neighborhood_search(SphHandle* h);
/* Host-side view of for_all_neighbors (base_container.py:549-560) for custom tasks and
 * tests: CSR neighbour lists of all particles in current order.  offsets has n+1 entries;
 * pass indices = NULL to query only the counts (offsets[n] = total). */
int sph_get_neighbors(SphHandle* h, int32_t* offsets, int32_t* indices, size_t indices_capacity);
/* per-cell particle counts in the reference's z-fastest flatten (base_container.py:472-481),
 * inclusive-scanned like grid_num_particles after PrefixSumExecutor.run (:546). */
int sph_get_grid_num_particles(SphHandle* h, int32_t* dst, size_t count);

/* ---- kernels ------------------------------------------------------------------------ */
int sph_run_task(SphHandle* h, int32_t task, int32_t iarg, float* out);

/* Whole solver step(s): BaseSolver.step (base_solver.py:692-696) around
 * {WCSPH.py:27-45, PCISPH.py:165-185, DFSPH.py:298-319}._step, for scenes whose rigid
 * solver step / late insert_object are no-ops (no dynamic rigid body, nothing pending);
 * the Python host falls back to task-by-task stepping otherwise. */
int sph_step(SphHandle* h, int32_t n_steps, SphStepStats* stats);
/* Solver loops alone (used by the task-by-task Python path to keep iteration logic native):
 * DFSPH.correct_density_error :225-243, correct_divergence_error :139-159,
 * PCISPH.refine :110-125, BaseSolver.implicit_viscosity_solve :509-517. */
int sph_dfsph_correct_density_error(SphHandle* h, int32_t* iterations, float* error);
int sph_dfsph_correct_divergence_error(SphHandle* h, int32_t* iterations, float* error);
int sph_pcisph_refine(SphHandle* h, int32_t* iterations, float* error);
int sph_implicit_viscosity_solve(SphHandle* h, int32_t* iterations, float* error);

int sph_synchronize(SphHandle* h);
/* Run all work of this handle on the caller's CUDA stream (cudaStream_t as void*; NULL restores the
 * handle's own stream) so that host frameworks can order / time it with their own events. */
int sph_set_stream(SphHandle* h, void* cuda_stream);

/* Per-kernel device time: when enabled every kernel launch is bracketed by CUDA events on the
 * handle's stream; sph_profile_read synchronises, returns one row per kernel name and resets. */
typedef struct SphKernelStat {
    char name[56];
    int64_t launches;
    double total_ms;
} SphKernelStat;
int sph_profile_enable(SphHandle* h, int32_t enable);
int sph_profile_read(SphHandle* h, SphKernelStat* out, int32_t capacity, int32_t* count);

/* ---- Z-slab sharding across the GPUs of one box (no reference counterpart) ----
 * One process + one handle (created with SPH_FLAG_SLAB) per GPU.  A slab owns the cell layers cz in
 * [z_lo, z_hi) of the GLOBAL grid and mirrors its neighbours' boundary layers z_lo-1 and z_hi as
 * read-only ghosts.  Migration, ghost import, per-field halo refreshes and the solver loops' error
 * sums run inside the library over NCCL (ncclSend/ncclRecv/ncclAllReduce on the handle's stream);
 * the host only brokers the NCCL unique id (e.g. with torch.distributed.broadcast).
 * Supported solvers: WCSPH and DFSPH with standard viscosity. */
#define SPH_SLAB_RECORD_WORDS 24   /* 32-bit words per migrating / ghost particle record */
typedef struct SphSlabInfo {
    int32_t z_lo, z_hi;           /* owned cell layers */
    int32_t n_owned;              /* particles with cz in [z_lo, z_hi) after the last sort */
    int32_t n_ghost;              /* imported ghosts */
    int32_t n_send_lo, n_send_hi; /* my boundary layers = the neighbours' ghost layers */
    int32_t own_begin, own_end;   /* owned index range of the sorted arrays */
    int64_t halo_bytes;           /* bytes sent so far (migration + ghosts + halo refreshes) */
    int64_t halo_calls;           /* halo refreshes so far */
} SphSlabInfo;
/* rank 0 creates the id (ncclGetUniqueId); every rank passes the same 128 bytes to sph_slab_init */
int sph_slab_unique_id(void* out128);
/* collective over all ranks: creates the communicator; add particles (owned ones only) afterwards.
 * global_particle_num = particle_num of the whole domain (DFSPH error normalisation, DFSPH.py:211,294) */
int sph_slab_init(SphHandle* h, int32_t rank, int32_t world, const void* unique_id128, int32_t z_lo, int32_t z_hi,
                  int64_t global_particle_num);
int sph_slab_set_global_particle_num(SphHandle* h, int64_t n);
int sph_slab_info(SphHandle* h, SphSlabInfo* out);
/* Peer memory over NVLink / NVSwitch for the solver loops (optional; without it the loops use NCCL): every rank
 * exports CUDA IPC handles of its velocity / payload arrays and of a small control block (sph_slab_peer_export,
 * SPH_SLAB_PEER_BLOB_BYTES bytes), the host gathers the blobs (e.g. torch.distributed.all_gather_object) and hands
 * every other rank's blob to sph_slab_peer_import.  Inside the DFSPH loops a rank then reads its ghosts' new
 * velocities / kappa straight from the neighbour's arrays (one copy kernel behind a flag the neighbour's sweep sets
 * in its epilogue) and the ranks exchange their error sums by storing them into each other's control blocks,
 * instead of two ncclSend/ncclRecv pairs and one ncclAllReduce per iteration. */
#define SPH_SLAB_PEER_BLOB_BYTES 512
int sph_slab_peer_export(SphHandle* h, void* blob /*[SPH_SLAB_PEER_BLOB_BYTES]*/);
int sph_slab_peer_import(SphHandle* h, int32_t peer_rank, const void* blob);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
