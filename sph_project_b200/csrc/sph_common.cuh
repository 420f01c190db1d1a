// sph_common.cuh — device state, math and the cell-window neighbour walk shared by all kernels.
//
// Data layout in HBM (all arrays cell-sorted, x-fastest / z-slowest flatten so that a Z-slab and
// its ghost layers are contiguous index ranges):
//   pv[i]  = float4(x, y, z, +-V)   position + rest volume; sign(w) < 0 marks a non-fluid particle,
//            so the hot loops read material and volume with the same 16-byte load as the position
//   vm[i]  = float4(vx, vy, vz, m)  velocity + mass
//   acc[i] = float4(ax, ay, az, 0)
//   scalar SoA: rho, p, alpha, kappa, ... (f32), material / object_id / is_dynamic / uid (i32)
//   cell_start[c] = exclusive scan of per-cell counts, ncell + 1 entries
// The 11 fields the reference permutes in reorder_particles (base_container.py:517-542) are
// double-buffered and gathered once per sort (ping-pong, no copy-back).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sph_b200.h"

#ifndef SPH_BLOCK   // variant builds may override it (Makefile `variants`); every kernel and chunk size follows
#define SPH_BLOCK 128
#endif

struct Consts {
    int N;            // particle_num
    int cap;          // particle_max_num
    int nx, ny, nz;   // grid_num
    int ncell;
    int has_dynamic_rigid;  // any particle with material rigid && is_dynamic
    float h;          // support radius = cell size
    float inv_h;
    float h2_thresh;  // r2 < h2_thresh  <=>  sqrtf(r2) < h  (exact, computed on the host)
    float kW, kW2, kG;  // 8/(pi h^3), 2x, 6x   (base_solver.py:56-103)
    float kG_inv_h;   // kG / h
    float V0, rho0, inv_rho0, dt, inv_dt, g_upper;
    float gx, gy, gz;
    float visc_cf, visc_cb;  // 2 (dim+2) mu, 2 (dim+2) mu_b  (base_solver.py:252,263)
    float visc_eps;   // 0.01 h^2
    float sigma;      // surface tension 0.01
    float diameter, diameter2, w_diameter;  // 2 dx, its square, W(2 dx)  (base_solver.py:222-229)
    float dom_x, dom_y, dom_z, padding;
    float pcisph_k;
    float corr_thresh;   // m_eps * dt of the DFSPH correction steps (DFSPH.py:175,258)
    int z_lo, z_hi;   // owned cell layers (whole grid unless the handle is a slab)
    int ghost_lo, ghost_hi;   // 1: a neighbour rank owns the layer z_lo - 1 / z_hi (its particles are ghosts here)
    int row_begin, row_end;   // rows the kernels update: [0, N) or, for a Z-slab, the owned index range
    int nbx, nby, nbz;        // brick grid: ceil(grid_num / SPH_BRICK_{X,Y,Z})  (sph_brick.cuh)
};

// per-particle work happens for owned rows only; ghost rows (Z-slabs) are read-only neighbours
#define SPH_ROW_OR_RETURN(c, i) if ((i) < (c).row_begin || (i) >= (c).row_end) return
#define SPH_IS_ROW(c, i) ((i) >= (c).row_begin && (i) < (c).row_end)

// device pointers; `cur` selects the live half of the ping-pong buffers
struct Dev {
    float4* pv;   float4* pv_alt;
    float4* vm;   float4* vm_alt;
    float* x0;    float* x0_alt;       // 3 floats per particle
    float* rho;   float* rho_alt;
    int* object_id;  int* object_id_alt;
    int* material;   int* material_alt;
    int* color;      int* color_alt;   // 3 ints per particle
    int* is_dynamic; int* is_dynamic_alt;
    int* grid_id;    int* grid_id_alt;
    int* uid;        int* uid_alt;
    int* ghost_slot; int* ghost_slot_alt;  // slabs only
    // not permuted by the sort (recomputed before use)
    float4* acc;
    float* p;
    float *alpha, *kappa, *kappa_v, *rho_star, *drho;       // DFSPH
    float4 *a_p, *v_pred, *x_pred;                           // PCISPH
    float4 *cg_p, *v_orig, *cg_Ap, *cg_x, *cg_b, *cg_r;      // implicit viscosity
    float* cg_dinv;                                          // 9 floats per particle
    // grid
    int* cell_count;   // ncell (+ used as rank scratch)
    int* cell_start;   // ncell + 1
    int* rank;         // cap
    int* perm;         // cap
    int* scan_tmp;     // block sums
    // object tables
    int* object_material;     // [20]
    int* rigid_is_dynamic;    // [20]
    float* rigid_state;       // [20][24]: com0(3) com(3) rot(9) vel(3) omega(3) pad(3)
    float* rigid_wrench;      // [20][6] force, torque
    // reductions
    double* red;              // small scratch for block reductions [64]
    // per-particle neighbour lists, valid while positions are frozen (sort .. next position update), row-major:
    // row i = nbr16[i * nbr_kmax ..]: word 0 = number of neighbours n of fluid particle i, words 1 .. n = the WINDOW
    // SLOTS (16 bit) of its neighbours inside the shared-memory window of the brick that owns i (sph_brick.cuh), in walk
    // order.  n = 0xffff marks a row without a list (more than nbr_kmax - 1 neighbours, a window beyond 16 bits, a
    // particle that left its sorted cell): such rows re-derive their neighbours by walking the 27 cells in global memory.
    unsigned short* nbr16;
    int nbr_kmax;             // multiple of 16 (one 256-bit load = 16 slots)
    // bricks (compact tiles of SPH_BRICK_X x Y x Z cells, one CTA each): flags set by the sort's gather, compacted list
    int* brick_flag;          // [nbricks] brick owns at least one fluid(-to-be) row
    int* brick_list;          // [nbricks] active brick ids, ascending
    unsigned short* row_order;   // [cap] per brick, stored at the sorted indices of its owned rows in flat order: flat owned index of
                                 // its k-th working row (fluid, owned by this rank), written by the list build
    int* brick_nf;            // [nbricks] by position in brick_list: working rows in row_order, -1 = no row list
    int* brick_ctl;           // [0] active count  [1] ticket  [2] CTAs finished  [3] max window slots seen  [4] windows above the smem budget
    // per-sweep scalar payload of neighbour j, staged next to pv_j: (s0, s1, rho_j, m_j) with (s0, s1) =
    // (kappa, kappa / rho) for the DFSPH correction steps, (p / rho^2, p) for the pressure force
    float4* aux;
};

// ---- small vector helpers -------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float4 a) { return make_float3(a.x, a.y, a.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// canonical squared distance, identical to oracle/sph_oracle.cpp dist2(): fma chain in x, y, z
__device__ __forceinline__ float dist2(float3 r) { return fmaf(r.z, r.z, fmaf(r.y, r.y, r.x * r.x)); }

// ---- SPH kernel (base_solver.py:56-103) -----------------------------------------------------
// W(q), q = r/h in [0, 1]; branch-free select between the two polynomial pieces
__device__ __forceinline__ float kernel_W_q(const Consts& c, float q) {
    float q2 = q * q;
    float a = c.kW * fmaf(6.0f * q2, q - 1.0f, 1.0f);   // k (6 q^3 - 6 q^2 + 1)
    float t = 1.0f - q;
    float b = c.kW2 * t * t * t;                         // 2 k (1 - q)^3
    float w = q <= 0.5f ? a : b;
    return q <= 1.0f ? w : 0.0f;
}
// gradient = R * kernel_gradient_scale(r2): zero for r <= 1e-5 (and beyond the support)
__device__ __forceinline__ float kernel_gradient_scale(const Consts& c, float r2) {
    float rinv;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(r2));   // r2 is far above the denormals wherever the result is used
    float r = r2 * rinv;
    float q = r * c.inv_h;
    float t = 1.0f - q;
    float f = q <= 0.5f ? q * fmaf(3.0f, q, -2.0f) : -t * t;   // k' q (3q - 2)  |  -k' (1-q)^2
    float s = c.kG_inv_h * f * rinv;                            // / (r h)
    return (r2 > 1e-10f && q <= 1.0f) ? s : 0.0f;
}

// ---- grid -------------------------------------------------------------------------------------
// pos_to_index (base_container.py:467-469): trunc(x / h) per axis with an IEEE f32 divide so the
// cell coordinates are bit-identical to the oracle's; clamped into the grid.
__device__ __forceinline__ int3 cell_of(const Consts& c, float x, float y, float z) {
    int cx = (int)(x / c.h), cy = (int)(y / c.h), cz = (int)(z / c.h);
    cx = min(max(cx, 0), c.nx - 1);
    cy = min(max(cy, 0), c.ny - 1);
    cz = min(max(cz, 0), c.nz - 1);
    return make_int3(cx, cy, cz);
}
__device__ __forceinline__ int flatten(const Consts& c, int3 g) { return (g.z * c.ny + g.y) * c.nx + g.x; }

// for_all_neighbors (base_container.py:549-560): walk the 27-cell window of particle i.  With the
// x-fastest flatten the three x-adjacent cells of a row are one contiguous run of the sorted
// arrays, so the window is 9 runs.  visit(j, pj, R, r2) is called for every j != i with |R| < h.
template <class Visit>
__device__ __forceinline__ void for_all_neighbors(const Consts& c, const Dev& d, int i, float4 pi, Visit&& visit) {
    const int3 g = cell_of(c, pi.x, pi.y, pi.z);
    const int xlo = max(g.x - 1, 0), xhi = min(g.x + 1, c.nx - 1);
#pragma unroll 1
    for (int dz = -1; dz <= 1; dz++) {
        const int zz = g.z + dz;
        if (zz < 0 || zz >= c.nz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            const int yy = g.y + dy;
            if (yy < 0 || yy >= c.ny) continue;
            const int row = (zz * c.ny + yy) * c.nx;
            const int s = __ldg(d.cell_start + row + xlo);
            const int e = __ldg(d.cell_start + row + xhi + 1);
#pragma unroll 1
            for (int j = s; j < e; j++) {
                const float4 pj = __ldg(d.pv + j);
                const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                const float r2 = dist2(R);
                if (r2 < c.h2_thresh && j != i) visit(j, pj, R, r2);
            }
        }
    }
}

// ---- host-side handle -------------------------------------------------------------------------
#include <string>
#include <vector>

struct SphHandle {
    SphParams P;
    Consts c;
    Dev d;
    cudaStream_t stream = nullptr;
    std::string err;
    int Nfluid = 0;
    bool sorted_valid = false;   // cell_start matches the current positions' sort
    std::vector<void*> allocations;
    // host mirrors
    float cg_alpha = 0, cg_beta = 0, cg_error = 0, density_error = 0;
    int32_t object_material[SPH_MAX_OBJECTS] = {0};
    int32_t rigid_is_dynamic_h[SPH_MAX_OBJECTS] = {0};
    float rigid_state_h[SPH_MAX_OBJECTS][24] = {{0}};
    // pinned scratch for scalar read-backs
    double* h_red = nullptr;
    void* staging = nullptr;       // device, cap * 36 bytes
    size_t staging_bytes = 0;
    int64_t launches = 0;
    int scan_blocks_cap = 0;
    bool dyn_rigid_dirty = true;
    bool lists_enabled = true;   // SPH_B200_NO_LISTS=1 forces window walks (A/B testing)
    bool list_valid = false;     // nbr lists match the current positions and order
    bool bricks_dirty = false;   // materials were edited by the host since the sort flagged the working bricks
    bool rigid_volume_clean = false;   // static boundary volumes are current (nothing added / edited since)
    // Z-slab state (sph_slab.cu); ghost_stale = fields whose ghost copies lag their owners
    struct SlabState* slab = nullptr;
    long long n_global = 0;
    int ghost_stale = 0;
    bool rows_from_sort = false; // owned range was set by a slab sort (host edits must not widen it to the ghosts)
    bool peer_loop = false;      // inside a DFSPH solve whose halos / error sums go through peer memory (sph_slab.cu)
    int peer_signalled = 0;      // ghost fields whose last writer signalled its completion to the peers
    int sticky_rc = 0;           // first error raised inside a void launcher (NCCL), reported by the caller
    int wmax = 2112;             // shared-memory window budget (slots per brick) of the sweep kernels
    int nbricks = 0;
    int solve_hint[2] = {2, 2};  // iterations the last DFSPH density / divergence solve took: size of the next first batch
    int solve_batch = 0;         // > 0: fixed number of solver iterations per host read (SPH_B200_BATCH_ITERS)
    // per-kernel event timing (sph_profile_enable / sph_profile_read)
    cudaStream_t own_stream = nullptr;
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t begin, end; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> event_pool;
};

// brackets one kernel launch with CUDA events when profiling is on
int sph_prof_begin(SphHandle* h, const char* name);
void sph_prof_end(SphHandle* h, int idx);
struct SphProf {
    SphHandle* h;
    int idx;
    SphProf(SphHandle* handle, const char* name) : h(handle), idx(-1) {
        if (h->profiling) idx = sph_prof_begin(h, name);
    }
    ~SphProf() {
        if (idx >= 0) sph_prof_end(h, idx);
    }
};
