// sph_sweeps.cu — per-particle neighbour summations ("sweeps") on brick tiles (sph_brick.cuh).
//
// Positions are frozen between a sort and the next position update and every solver runs many sweeps in that
// interval (DFSPH: density, alpha, the divergence solve, then next step's surface tension, viscosity and the
// ~25-iteration density solve).  So:
//   * the first sweep after every sort (compute_density) builds neighbour lists: one warp per cell tests the ~216
//     candidates of the cell's 27-cell neighbourhood — read once from the brick's TMA-staged window, 32 candidates
//     per step across the lanes — against each particle of the cell and appends the accepted window slots to that
//     particle's row with a ballot + prefix-popcount compaction (coalesced 16-bit stores, walk order preserved);
//   * every sweep, that one included, then runs one thread per owned fluid particle: it streams its row (16 slots
//     per 256-bit load) and reads neighbour j's position and payload from the brick's shared-memory window;
//   * rows without a list, and everything when lists are disabled (SPH_B200_NO_LISTS=1), re-derive their
//     neighbours by walking the 27 cells in global memory (same order, same arithmetic, same result bit for bit).
//
// Each kernel replaces one @ti.kernel + its *_task of the reference (cited per kernel); the task bodies are inlined
// lambdas where upstream passes ti.template() callbacks into for_all_neighbors (base_container.py:549-560).
#include <map>
#include <utility>

#include "sph_kernels.h"
#include "sph_brick.cuh"

namespace {

extern __shared__ __align__(16) float4 dyn_smem[];

// resident CTAs per SM the register allocator has to make room for, by number of staged arrays (the window budget
// of 2112 slots x 16 B x arrays allows 4 / 3 / 2 CTAs of 512 threads per SM)
#ifndef SPH_BRICK_MINB1
#define SPH_BRICK_MINB1 3
#endif
#ifndef SPH_BRICK_MINB2
#define SPH_BRICK_MINB2 3
#endif
#ifndef SPH_BRICK_MINB3
#define SPH_BRICK_MINB3 2
#endif

// rigid_body_forces / rigid_body_torques accumulation (base_solver.py:174-187 and twins)
__device__ __forceinline__ void add_wrench(const Dev& d, int obj, float3 force, float3 at) {
    if (obj < 0 || obj >= SPH_MAX_OBJECTS) return;
    const float* st = d.rigid_state + obj * 24;
    float3 arm = make_float3(at.x - st[3], at.y - st[4], at.z - st[5]);
    float3 tq = cross3(arm, force);
    float* w = d.rigid_wrench + obj * 6;
    atomicAdd(w + 0, force.x); atomicAdd(w + 1, force.y); atomicAdd(w + 2, force.z);
    atomicAdd(w + 3, tq.x); atomicAdd(w + 4, tq.y); atomicAdd(w + 5, tq.z);
}

// Rows of the open brick.  With the build's compacted row list (sh.nf >= 0): warps draw groups of 32 consecutive
// working rows from the brick's shared counter until none is left (no lanes parked on boundary particles, no warp
// waiting while another one still has two groups to go); otherwise every owned row in flat order, one per thread and pass.
template <class RowFn>
__device__ __forceinline__ void brick_rows(const Consts& c, const Dev& d, const Brick& bk, RowFn&& row) {
    BrickShared& sh = *bk.sh;
    const int nf = sh.nf;
    if (nf >= 0) {
        const int lane = threadIdx.x & 31;
        const int groups = (nf + 31) >> 5;
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(&sh.group, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= groups) break;
            const int ts = g * 32 + lane;
            if (ts < nf) {
                const int t = d.row_order[brick_own_row(bk, ts).i];
                row(bk, brick_own_row(bk, t));   // only working rows of this rank are in the order
            }
        }
    } else {
        const int nown = brick_own_count(bk);
        for (int t = threadIdx.x; t < nown; t += SPH_BRICK_THREADS) {
            const BrickRow r = brick_own_row(bk, t);
            if (SPH_IS_ROW(c, r.i)) row(bk, r);
        }
    }
}

// Persistent CTA: draw bricks, stage the window of NARR arrays, call row(bk, r) for every owned row of the brick.
// lists: the neighbour lists (and with them the sorted row order) are valid.
template <int NARR, class RowFn>
__device__ __forceinline__ void brick_for_rows(const Consts& c, const Dev& d, int wmax, bool lists, const float4* g0, const float4* g1,
                                               const float4* g2, RowFn&& row, const PeerLinks* peer = nullptr) {
    __shared__ BrickSmem bsm;
    Brick bk;
    brick_begin(c, d, bk, &bsm, dyn_smem, wmax, g0, g1, g2, lists, lists, peer);
    while (brick_stage<NARR>(c, d, bk)) {
        brick_rows(c, d, bk, row);
        brick_advance(bk);
    }
}

// aux[i] = (s0, s1, rho_i, m_i), the scalar payload of the pressure / correction / viscosity sweeps
enum AuxMode { AUX_RHO_M, AUX_KAPPA, AUX_KAPPA_V, AUX_PRESSURE };
template <int MODE>
__global__ void __launch_bounds__(SPH_BLOCK) k_prep_aux(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    const float rho = d.rho[i];
    const float m = reinterpret_cast<const float*>(d.vm + i)[3];
    float s0 = 0.f, s1 = 0.f;
    if (MODE == AUX_KAPPA || MODE == AUX_KAPPA_V) {
        s0 = MODE == AUX_KAPPA ? d.kappa[i] : d.kappa_v[i];
        s1 = s0 / rho;
    } else if (MODE == AUX_PRESSURE) {
        s1 = d.p[i];
        s0 = s1 / (rho * rho);
    }
    d.aux[i] = make_float4(s0, s1, rho, m);
}

// compute_rigid_particle_volume (base_solver.py:105-123); rigid rows, plain window walk in global
// memory (once per step over the boundary shell, skipped while the boundaries are static)
__global__ void __launch_bounds__(SPH_BLOCK) k_rigid_volume(Consts c, Dev d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    float4 pi = d.pv[i];
    if (!(pi.w < 0.0f) || !(pi.y <= c.g_upper)) return;
    const int obj_i = d.object_id[i];
    float ret = c.kW;  // W(0)
    for_all_neighbors(c, d, i, pi, [&](int j, float4, float3, float r2) {
        if (__ldg(d.object_id + j) == obj_i) ret += kernel_W_q(c, sqrtf(r2) * c.inv_h);
    });
    const float V = 1.0f / ret, m = c.rho0 * V;
    reinterpret_cast<float*>(d.pv + i)[3] = -V;
    reinterpret_cast<float*>(d.vm + i)[3] = m;
}

// ---- list build + compute_density (+ DFSPH compute_alpha) ------------------------------------------------
// Neighbour list of one owned fluid particle: its 9 candidate runs (three x-adjacent cells each, walk order) are
// contiguous slot ranges of the brick's window; every candidate is read from shared memory (lanes of one cell read
// the same address: broadcast) and the accepted slots are appended to the particle's row.
__device__ __forceinline__ void brick_build_row(const Consts& c, const Dev& d, const Brick& bk, BrickRow row, float4 pi) {
    const BrickShared& sh = *bk.sh;
    const int kcap = d.nbr_kmax - 1;
    unsigned short* out = d.nbr16 + (size_t)row.i * d.nbr_kmax;
    // the particle's sorted cell inside the brick
    const int r_own = row.run;
    int lx = 0;
#pragma unroll
    for (int k = 2; k <= BRK_X; k++) lx += (row.i >= sh.T[r_own][k]) ? 1 : 0;
    const int ly = r_own % BRK_RY - 1, lz = r_own / BRK_RY - 1;
    // a particle that left its sorted cell since the sort (host edits, sweeps on a stale grid) searches around its
    // CURRENT cell like the reference does: no list, global walk; same for windows beyond 16 bits
    const int3 cc = cell_of(c, pi.x, pi.y, pi.z);
    if (!(cc.x == sh.x0 + lx && cc.y == sh.y0 + ly && cc.z == sh.z0 + lz) || sh.S[BRK_RUNS] > 65535) {
        out[0] = (unsigned short)SPH_ROW_NO_LIST;
        return;
    }
    const bool staged = sh.staged != 0;
    unsigned a0 = bk.a0;
    asm volatile("mov.u32 %0, %0;" : "+r"(a0));
    const float h2 = c.h2_thresh;
    int n = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const int r = (lz + k / 3) * BRK_RY + (ly + k % 3);   // run of (lz + dz, ly + dy), dz = k / 3 - 1, dy = k % 3 - 1
        const int g = sh.T[r][lx];
        const int len = sh.T[r][lx + 3] - g;
        int slot = sh.S[r] + (g - sh.T[r][0]);
        if (staged) {
#pragma unroll 2
            for (int e = 0; e < len; e++, slot++) {
                const float4 pj = lds128(a0 + 16u * (unsigned)slot);
                const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                const bool acc = dist2(R) < h2 && slot != row.slot;
                if (acc && n < kcap) out[1 + n] = (unsigned short)slot;
                n += acc ? 1 : 0;
            }
        } else {
#pragma unroll 1
            for (int e = 0; e < len; e++, slot++) {
                const float4 pj = __ldg(d.pv + g + e);
                const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                const bool acc = dist2(R) < h2 && slot != row.slot;
                if (acc && n < kcap) out[1 + n] = (unsigned short)slot;
                n += acc ? 1 : 0;
            }
        }
    }
    out[0] = n <= kcap ? (unsigned short)n : (unsigned short)SPH_ROW_NO_LIST;   // too many neighbours: the row walks
}

// compute_density (base_solver.py:521-541) of one row; visitor shared by the list and the walk paths
template <bool LIST, bool NC>
__device__ __forceinline__ void row_density(const Consts& c, const Dev& d, const Brick& bk, int i, float4 pi) {
    float ret = 0.0f;
    brick_neighbors<LIST, NC, 1>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4, float4, float3, float r2) {
        ret = fmaf(fabsf(pj.w), kernel_W_q(c, sqrtf(r2) * c.inv_h), ret);
    });
    d.rho[i] = fmaf(pi.w, c.kW, ret) * c.rho0;
}

// DFSPH compute_alpha (DFSPH.py:22-62) of one row
template <bool LIST, bool NC>
__device__ __forceinline__ void row_alpha(const Consts& c, const Dev& d, const Brick& bk, int i, float4 pi) {
    float3 grad_i = make_float3(0.f, 0.f, 0.f);
    float sum_k = 0.0f;
    brick_neighbors<LIST, NC, 1>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4, float4, float3 R, float r2) {
        const float s = -fabsf(pj.w) * kernel_gradient_scale(c, r2);
        const float3 g = R * s;
        if (pj.w > 0.0f) sum_k += dist2(g);
        grad_i = grad_i + g;
    });
    sum_k += dist2(grad_i);
    d.alpha[i] = sum_k > 1e-5f ? 1.0f / sum_k : 0.0f;
}

// BUILD: neighbour lists of every active brick and the compacted list of its working rows (fluid rows of this rank in
// flat owned order: a stable compaction, so consecutive lanes stay spatial neighbours and share window slots);
// DENSITY / ALPHA: the two position-only sweeps that follow a sort, on the window the build already staged.
template <bool BUILD, bool DENSITY, bool ALPHA>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB1) kb_build(Consts c, Dev d, int wmax) {
    __shared__ BrickSmem bsm;
    __shared__ int s_warp[BRK_SORT_PASSES * BRK_WARPS + 1];   // working rows per (pass, warp), then exclusive offsets
    Brick bk;
    brick_begin(c, d, bk, &bsm, dyn_smem, wmax, d.pv, nullptr, nullptr, BUILD, false);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (brick_stage<1>(c, d, bk)) {
        if (BUILD) {
            const int nown = brick_own_count(bk);
            const bool compactable = nown <= BRK_ROWS_MAX;
            int rank[BRK_SORT_PASSES];
            bool works[BRK_SORT_PASSES];
            for (int t0 = 0; t0 < nown; t0 += BRK_ROWS_MAX) {   // one trip unless the brick is absurdly full
#pragma unroll
                for (int p = 0; p < BRK_SORT_PASSES; p++) {
                    const int t = t0 + p * SPH_BRICK_THREADS + (int)threadIdx.x;
                    bool w = false;
                    if (t < nown) {
                        const BrickRow row = brick_own_row(bk, t);
                        if (SPH_IS_ROW(c, row.i)) {
                            const float4 pi = brick_own_load(bk, row, 0);
                            if (pi.w > 0.0f) {
                                brick_build_row(c, d, bk, row, pi);
                                w = true;
                            }
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, w);
                    works[p] = w;
                    rank[p] = __popc(m & ((1u << lane) - 1u));
                    if (lane == 0) s_warp[p * BRK_WARPS + warp] = __popc(m);
                }
            }
            __syncthreads();
            if (threadIdx.x < 32) {   // exclusive scan over (pass, warp): flat owned order
                int carry = 0;
#pragma unroll
                for (int e0 = 0; e0 < BRK_SORT_PASSES * BRK_WARPS; e0 += 32) {
                    const int e = e0 + lane;
                    const int v = e < BRK_SORT_PASSES * BRK_WARPS ? s_warp[e] : 0;
                    const int inc = brk_warp_inclusive_scan(v);
                    if (e < BRK_SORT_PASSES * BRK_WARPS) s_warp[e] = carry + inc - v;
                    carry += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (lane == 0) s_warp[BRK_SORT_PASSES * BRK_WARPS] = carry;
            }
            __syncthreads();
            if (compactable) {
#pragma unroll
                for (int p = 0; p < BRK_SORT_PASSES; p++) {
                    const int t = p * SPH_BRICK_THREADS + (int)threadIdx.x;
                    // the brick's flat owned positions double as the storage of its row list: disjoint across bricks
                    if (works[p]) d.row_order[brick_own_row(bk, s_warp[p * BRK_WARPS + warp] + rank[p]).i] = (unsigned short)t;
                }
            }
            if (threadIdx.x == 0) {
                const int nf = compactable ? s_warp[BRK_SORT_PASSES * BRK_WARPS] : -1;
                d.brick_nf[bk.sh->ordinal] = nf;
                bk.sh->nf = nf;
            }
            __syncthreads();   // lists and row list are visible to the whole CTA
        }
        if (DENSITY || ALPHA) {
            brick_rows(c, d, bk, [&](const Brick& bk_, BrickRow row) {
                const float4 pi = brick_own_load(bk_, row, 0);
                if (!(pi.w > 0.0f)) return;
                if (DENSITY) row_density<BUILD, false>(c, d, bk_, row.i, pi);
                if (ALPHA) row_alpha<BUILD, false>(c, d, bk_, row.i, pi);
            });
        }
        brick_advance(bk);
    }
    brick_finish(d);
}

// position-only sweeps on existing lists
template <bool LIST, bool DENSITY>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB1) kb_pv_sweep(Consts c, Dev d, int wmax) {
    brick_for_rows<1>(c, d, wmax, LIST, d.pv, nullptr, nullptr, [&](const Brick& bk, BrickRow row) {
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        if (DENSITY) row_density<LIST, true>(c, d, bk, row.i, pi);
        else row_alpha<LIST, true>(c, d, bk, row.i, pi);
    });
    brick_finish(d);
}

// compute_pressure_acceleration (base_solver.py:135-187) and, with TEMP, PCISPH's
// compute_temp_pressure_acceleration (PCISPH.py:74-107: fluid rows, no rigid wrench, output a_p).
// aux = (p / rho^2, p, rho, m).  Rows outside the active bricks hold the zeros of the launcher's memset.
template <bool TEMP, bool LIST>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB2) kb_pressure_accel(Consts c, Dev d, int wmax) {
    brick_for_rows<2>(c, d, wmax, LIST, d.pv, d.aux, nullptr, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        float4* out = TEMP ? d.a_p : d.acc;
        bool active = pi.w > 0.0f;
        if (!TEMP) active = active && d.is_dynamic[i] != 0;
        if (!active) {
            out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
        const float dpi = brick_own_load(bk, row, 1).x;
        float3 ret = make_float3(0.f, 0.f, 0.f);
        brick_neighbors<LIST, true, 2>(c, d, bk, i, pi, [&](NbrRef ref, float4 pj, float4 aj, float4, float3 R, float r2) {
            const float gs = kernel_gradient_scale(c, r2);
            float coef;
            if (pj.w > 0.0f) {
                coef = -aj.w * (dpi + aj.x);
            } else {
                coef = -c.rho0 * (-pj.w) * dpi;
                if (!TEMP && c.has_dynamic_rigid) {
                    const int j = brick_nbr_index(bk, ref);
                    if (__ldg(d.is_dynamic + j)) {
                        float3 force = R * (-coef * gs * (c.rho0 * pi.w));
                        add_wrench(d, __ldg(d.object_id + j), force, f3(pi));  // arm from x_i (base_solver.py:185)
                    }
                }
            }
            const float s = coef * gs;
            ret.x = fmaf(s, R.x, ret.x); ret.y = fmaf(s, R.y, ret.y); ret.z = fmaf(s, R.z, ret.z);
        });
        out[i] = make_float4(ret.x, ret.y, ret.z, 0.f);
    });
    brick_finish(d);
}

// compute_surface_tension_acceleration (base_solver.py:209-229): a_i += sum.  Payload vm: m_j in w
template <bool LIST>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB2) kb_surface_tension(Consts c, Dev d, int wmax) {
    brick_for_rows<2>(c, d, wmax, LIST, d.pv, d.vm, nullptr, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float sm = c.sigma / brick_own_load(bk, row, 1).w;
        float3 a = make_float3(0.f, 0.f, 0.f);
        brick_neighbors<LIST, true, 2>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4 vj, float4, float3 R, float r2) {
            if (!(pj.w > 0.0f)) return;
            const float w = r2 > c.diameter2 ? kernel_W_q(c, sqrtf(r2) * c.inv_h) : c.w_diameter;
            const float s = sm * vj.w * w;
            a.x = fmaf(-s, R.x, a.x); a.y = fmaf(-s, R.y, a.y); a.z = fmaf(-s, R.z, a.z);
        });
        float4 acc = d.acc[i];
        d.acc[i] = make_float4(acc.x + a.x, acc.y + a.y, acc.z + a.z, 0.f);
    });
    brick_finish(d);
}

// compute_viscosity_acceleration_standard (base_solver.py:231-278): a_i += sum / rho0.
// Payloads vm = (v_j, m_j) and aux = (., ., rho_j, .)
template <bool LIST>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB3) kb_viscosity(Consts c, Dev d, int wmax) {
    brick_for_rows<3>(c, d, wmax, LIST, d.pv, d.vm, d.aux, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float4 vi = brick_own_load(bk, row, 1);
        const float den_i = brick_own_load(bk, row, 2).z;
        float3 a = make_float3(0.f, 0.f, 0.f);
        brick_neighbors<LIST, true, 3>(c, d, bk, i, pi, [&](NbrRef ref, float4 pj, float4 vj, float4 xj, float3 R, float r2) {
            const float v_xy = dot3(make_float3(vi.x - vj.x, vi.y - vj.y, vi.z - vj.z), R);
            const float gs = kernel_gradient_scale(c, r2);
            float coef;
            if (pj.w > 0.0f) {
                coef = c.visc_cf * ((vi.w + vj.w) * 0.5f) / xj.z;
            } else {
                coef = c.visc_cb * (c.rho0 * (-pj.w)) / den_i;
            }
            const float s = coef / (r2 + c.visc_eps) * v_xy * gs;
            a.x = fmaf(s, R.x, a.x); a.y = fmaf(s, R.y, a.y); a.z = fmaf(s, R.z, a.z);
            if (!(pj.w > 0.0f) && c.has_dynamic_rigid) {
                const int j = brick_nbr_index(bk, ref);
                if (__ldg(d.is_dynamic + j)) {
                    float3 force = R * (-s * vi.w * c.inv_rho0);     // -acc * m_i / rho0
                    add_wrench(d, __ldg(d.object_id + j), force, f3(pj));
                }
            }
        });
        float4 acc = d.acc[i];
        d.acc[i] = make_float4(fmaf(a.x, c.inv_rho0, acc.x), fmaf(a.y, c.inv_rho0, acc.y), fmaf(a.z, c.inv_rho0, acc.z), 0.f);
    });
    brick_finish(d);
}

// Loop control of the DFSPH solves when the exit test runs on the device (SolveCtl::mode):
//   SOLVE_PLAIN   kernels of the task API: no error sum, no test
//   SOLVE_SUM     add the row errors into red[RED_ERR]; the host (or, for Z-slabs, the all-reduce) takes it from there
//   SOLVE_TEST    as SOLVE_SUM, then the last CTA of the grid applies the reference's loop-exit test (DFSPH.py:152-157,
//                 236-242): iterations += 1, error = sum / particle_num, done = error <= eta; later launches of the batch
//                 see `done` and return at once, so the solve stops at the same iteration as a host-driven loop
enum SolveMode { SOLVE_PLAIN = 0, SOLVE_SUM = 1, SOLVE_TEST = 2 };
struct SolveCtl {
    int mode;
    int speculative;   // launched ahead of the exit test: return at once when red[CTRL_DONE] is set
    float n_global;    // particle_num of the whole domain (the reference averages over fluid AND boundary rows)
    float eta;
    PeerLinks peer;    // Z-slab peer loops: whom to tell that this sweep is complete / where the error sum goes
};

// Epilogue of a sweep inside a peer loop (thread 0 of the last CTA to finish; every write of the grid is visible here):
// count the completed writer where the neighbours' ghost pulls look for it.
__device__ __forceinline__ void peer_signal(const PeerLinks& p, bool aux) {
    if (!p.mine) return;
    __threadfence_system();
    int* cnt = aux ? &p.mine->aux_count : &p.mine->vel_count;
    *(volatile int*)cnt = *(volatile int*)cnt + 1;
}
// ... and deliver this rank's error sum into every rank's control block (own included), then clear it
__device__ __forceinline__ void peer_deliver_error(const PeerLinks& p, double* red_err) {
    const double part = *(volatile double*)red_err;
    const int count = *(volatile int*)&p.mine->err_count + 1;
    for (int q = 0; q < p.world; q++) *(volatile double*)&p.all[q]->partial[count & 1][p.rank] = part;
    __threadfence_system();
    for (int q = 0; q < p.world; q++) *(volatile int*)&p.all[q]->err_flag[p.rank] = count;
    *(volatile int*)&p.mine->err_count = count;
    *red_err = 0.0;
}

// DFSPH compute_density_derivative (DFSPH.py:65-101) / compute_density_star (:104-126).
// FUSED (the library's own solver loops): also the kappa of the next correction step
// (compute_kappa_v :132-137 / compute_kappa :217-223, written to the field and to aux) and the
// error sum (compute_density_derivative_error :205-211 / compute_density_error :285-294).
template <bool STAR, bool LIST, bool FUSED>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB2) kb_dfsph_density_change(Consts c, Dev d, int wmax, SolveCtl ctl) {
    if (ctl.speculative && d.red[CTRL_DONE] != 0.0) return;   // uniform over the grid
    float err = 0.0f;
    brick_for_rows<2>(c, d, wmax, LIST, d.pv, d.vm, nullptr, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float4 vi = brick_own_load(bk, row, 1);
        float delta = 0.0f;
        const int nn = brick_neighbors<LIST, true, 2>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4 vj, float4, float3 R, float r2) {
            const float vr = dot3(make_float3(vi.x - vj.x, vi.y - vj.y, vi.z - vj.z), R);
            delta = fmaf(fabsf(pj.w) * kernel_gradient_scale(c, r2), vr, delta);
        });
        const float rho = d.rho[i];
        float kap;
        if (STAR) {
            const float rs = fmaxf(fmaf(c.dt, delta, rho / c.rho0), 1.0f);
            d.rho_star[i] = rs;
            kap = (rs - 1.0f) * d.alpha[i] * c.inv_dt;
            if (FUSED) { d.kappa[i] = kap; err += rs - 1.0f; }
        } else {
            float adv = fmaxf(delta, 0.0f);
            if (nn < 20) adv = 0.0f;   // particle deficiency (DFSPH.py:93-95)
            d.drho[i] = adv;
            kap = adv * d.alpha[i];
            if (FUSED) { d.kappa_v[i] = kap; err += c.rho0 * adv; }
        }
        if (FUSED) d.aux[i] = make_float4(kap, kap / rho, rho, vi.w);
    }, &ctl.peer);
    if (FUSED && ctl.mode != SOLVE_PLAIN) block_reduce_add(d.red + RED_ERR, (double)err);
    const bool last = brick_finish(d);
    if (FUSED && last) {
        peer_signal(ctl.peer, true);   // aux (kappa) of my boundary layers is final
        if (ctl.peer.mine && ctl.mode == SOLVE_SUM) peer_deliver_error(ctl.peer, d.red + RED_ERR);
    }
    if (FUSED && ctl.mode == SOLVE_TEST && last) {
        // host arithmetic: f32 division of the f64 sum by the particle count, compare with eta
        const float e = (float)(*(volatile double*)(d.red + RED_ERR)) / ctl.n_global;
        d.red[CTRL_ITERS] += 1.0;
        d.red[CTRL_ERR] = (double)e;
        if (e <= ctl.eta) d.red[CTRL_DONE] = 1.0;
        d.red[RED_ERR] = 0.0;
    }
}

// The same loop-exit test as a kernel of its own, for Z-slabs: red[RED_ERR] has just been summed over the ranks by the
// all-reduce, so every rank takes the same decision from the same number.
__global__ void k_dfsph_solve_check(Dev d, SolveCtl ctl) {
    if (d.red[CTRL_DONE] == 0.0) {
        if (ctl.peer.mine) {   // peer loop: wait for every rank's delivery of this round, sum in rank order
            const PeerCtl* me = ctl.peer.mine;
            const int want = *(volatile const int*)&me->err_count;
            double sum = 0.0;
            for (int q = 0; q < ctl.peer.world; q++) {
                while (*(volatile const int*)&me->err_flag[q] < want) __nanosleep(64);
            }
            __threadfence_system();
            for (int q = 0; q < ctl.peer.world; q++) sum += *(volatile const double*)&me->partial[want & 1][q];
            d.red[RED_ERR] = sum;
        }
        const float e = (float)d.red[RED_ERR] / ctl.n_global;
        d.red[CTRL_ITERS] += 1.0;
        d.red[CTRL_ERR] = (double)e;
        if (e <= ctl.eta) d.red[CTRL_DONE] = 1.0;
    }
    d.red[RED_ERR] = 0.0;
}

// DFSPH correct_divergence_step (DFSPH.py:161-202) / correct_density_error_step (:245-283).
// aux = (kappa_j, kappa_j / rho_j, rho_j, m_j); the new velocity goes to vm.  WRENCH: the scene holds dynamic rigid
// particles (force / torque accumulation on rigid neighbours).
template <bool LIST, bool WRENCH>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB2) kb_dfsph_correct(Consts c, Dev d, int wmax, SolveCtl ctl) {
    if (ctl.speculative && d.red[CTRL_DONE] != 0.0) return;   // the solve converged earlier in this batch
    brick_for_rows<2>(c, d, wmax, LIST, d.pv, d.aux, nullptr, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float4 ai = brick_own_load(bk, row, 1);
        const float k_i = ai.x, ki_rho = ai.y;
        float3 dv = make_float3(0.f, 0.f, 0.f);
        brick_neighbors<LIST, true, 2, true>(c, d, bk, i, pi, [&](NbrRef ref, float4 pj, float4 aj, float4, float3 R, float r2) {
            // fluid j: V_j gradW (kappa_i / rho_i + kappa_j / rho_j) rho0 unless |kappa_i + kappa_j| is below the threshold;
            // rigid j: V_j gradW kappa_i / rho_i rho0 unless |kappa_i| is.  Same products in the same order on both
            // branches; a pair below the threshold contributes s = 0 (dv - 0 R = dv exactly) instead of branching.
            const bool fluid = pj.w > 0.0f;
            const bool keep = fabsf(fluid ? k_i + aj.x : k_i) > c.corr_thresh;
            float s = fabsf(pj.w) * kernel_gradient_scale(c, r2) * (fluid ? ki_rho + aj.y : ki_rho) * c.rho0;
            s = keep ? s : 0.0f;
            if (WRENCH && keep && !fluid) {
                const int j = brick_nbr_index(bk, ref);
                if (__ldg(d.is_dynamic + j)) {
                    float3 force = R * (s * c.inv_dt * (pi.w * c.rho0));
                    add_wrench(d, __ldg(d.object_id + j), force, f3(pj));
                }
            }
            dv.x = fmaf(-s, R.x, dv.x); dv.y = fmaf(-s, R.y, dv.y); dv.z = fmaf(-s, R.z, dv.z);
        });
        float4 v = d.vm[i];
        d.vm[i] = make_float4(v.x + dv.x, v.y + dv.y, v.z + dv.z, v.w);
    }, &ctl.peer);
    if (brick_finish(d)) peer_signal(ctl.peer, false);   // velocities of my boundary layers are final
}

// PCISPH compute_density_star (PCISPH.py:32-62): predicted positions, no self term, neighbour
// set from the current positions.  Accumulates sum max(0, rho*/rho0 - 1) into red[RED_ERR].
template <bool LIST>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB2) kb_pcisph_density_star(Consts c, Dev d, int wmax) {
    float err = 0.0f;
    brick_for_rows<2>(c, d, wmax, LIST, d.pv, d.x_pred, nullptr, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float4 xi = brick_own_load(bk, row, 1);
        float ret = 0.0f;
        brick_neighbors<LIST, true, 2>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4 xpj, float4, float3, float) {
            const float4 xj = pj.w > 0.0f ? xpj : pj;
            const float r2 = dist2(make_float3(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z));
            ret = fmaf(fabsf(pj.w), kernel_W_q(c, sqrtf(r2) * c.inv_h), ret);
        });
        d.rho_star[i] = ret * c.rho0;
        err += fmaxf(0.0f, ret - 1.0f);
    });
    block_reduce_add(d.red + RED_ERR, (double)err);
    brick_finish(d);
}

// implicit viscosity: A_ij = -c (grad W_ij (x) R) / (r^2 + 0.01 h^2)  (base_solver.py:348-371);
// returns c' such that A_ij = c' * (R (x) R)   (grad W = gs * R)
__device__ __forceinline__ float visc_A_scale(const Consts& c, float mi, float den_i, float4 pj, float mj, float rho_j, float r2) {
    const float gs = kernel_gradient_scale(c, r2);
    float coef;
    if (pj.w > 0.0f) coef = -c.visc_cf * ((mi + mj) * 0.5f) / rho_j;
    else coef = -c.visc_cb * (c.rho0 * (-pj.w)) / den_i;
    return coef / (r2 + c.visc_eps) * gs;
}

// prepare_conjugate_gradient_solver1, the per-particle part (base_solver.py:300-315):
// D_i^-1, b_i and p_i <- x_i.  Payloads vm = (v_j, m_j), aux = (., ., rho_j, .)
template <bool LIST>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB3) kb_cg_prepare1(Consts c, Dev d, int wmax) {
    brick_for_rows<3>(c, d, wmax, LIST, d.pv, d.vm, d.aux, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float4 vi = brick_own_load(bk, row, 1);
        const float den_i = brick_own_load(bk, row, 2).z;
        // ret = -sum A_ij (symmetric in R (x) R): 6 unique entries
        float sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
        float3 b = make_float3(0.f, 0.f, 0.f);
        brick_neighbors<LIST, true, 3>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4 vj, float4 xj, float3 R, float r2) {
            const float rho_j = pj.w > 0.0f ? xj.z : 1.0f;
            const float a = -visc_A_scale(c, vi.w, den_i, pj, vj.w, rho_j, r2);   // ret -= A_ij
            sxx = fmaf(a * R.x, R.x, sxx); sxy = fmaf(a * R.x, R.y, sxy); sxz = fmaf(a * R.x, R.z, sxz);
            syy = fmaf(a * R.y, R.y, syy); syz = fmaf(a * R.y, R.z, syz); szz = fmaf(a * R.z, R.z, szz);
            if (!(pj.w > 0.0f)) {   // compute_b_i_task :333-346, rigid neighbours only
                const float s = c.visc_cb * c.rho0 * (-pj.w) / den_i * dot3(f3(vj), R) / (r2 + c.visc_eps) *
                                kernel_gradient_scale(c, r2);
                b.x = fmaf(s, R.x, b.x); b.y = fmaf(s, R.y, b.y); b.z = fmaf(s, R.z, b.z);
            }
        });
        // diag = I - ret * dt / rho0
        const float f = c.dt * c.inv_rho0;
        const float m00 = fmaf(-sxx, f, 1.0f), m01 = -sxy * f, m02 = -sxz * f, m11 = fmaf(-syy, f, 1.0f), m12 = -syz * f, m22 = fmaf(-szz, f, 1.0f);
        const float c00 = fmaf(m11, m22, -(m12 * m12)), c01 = fmaf(m12, m02, -(m01 * m22)), c02 = fmaf(m01, m12, -(m11 * m02));
        const float inv = 1.0f / fmaf(m02, c02, fmaf(m01, c01, m00 * c00));
        float* o = d.cg_dinv + 9 * (size_t)i;
        o[0] = c00 * inv; o[1] = c01 * inv; o[2] = c02 * inv;
        o[3] = c01 * inv; o[4] = fmaf(m00, m22, -(m02 * m02)) * inv; o[5] = fmaf(m02, m01, -(m00 * m12)) * inv;
        o[6] = c02 * inv; o[7] = o[5]; o[8] = fmaf(m00, m11, -(m01 * m01)) * inv;
        const float g = c.dt * c.inv_rho0;
        d.cg_b[i] = make_float4(fmaf(-g, b.x, vi.x), fmaf(-g, b.y, vi.y), fmaf(-g, b.z, vi.z), 0.f);
        d.cg_p[i] = d.cg_x[i];
    });
    brick_finish(d);
}

// compute_Ap (base_solver.py:373-391): Ap_i = p_i + dt/rho0 * D_i^-1 sum_{fluid j} (-A_ij) p_j.
// Payloads aux = (., ., rho_j, m_j) and cg_p
template <bool LIST>
__global__ void __launch_bounds__(SPH_BRICK_THREADS, SPH_BRICK_MINB3) kb_cg_Ap(Consts c, Dev d, int wmax) {
    brick_for_rows<3>(c, d, wmax, LIST, d.pv, d.aux, d.cg_p, [&](const Brick& bk, BrickRow row) {
        const int i = row.i;
        const float4 pi = brick_own_load(bk, row, 0);
        if (!(pi.w > 0.0f)) return;
        const float4 ai = brick_own_load(bk, row, 1);
        const float mi = ai.w, den_i = ai.z;
        float3 s = make_float3(0.f, 0.f, 0.f);
        brick_neighbors<LIST, true, 3>(c, d, bk, i, pi, [&](NbrRef, float4 pj, float4 aj, float4 pj_cg, float3 R, float r2) {
            if (!(pj.w > 0.0f)) return;
            const float a = -visc_A_scale(c, mi, den_i, pj, aj.w, aj.z, r2);
            const float t = a * dot3(R, f3(pj_cg));   // (-A_ij) p_j = a R (R . p_j)
            s.x = fmaf(t, R.x, s.x); s.y = fmaf(t, R.y, s.y); s.z = fmaf(t, R.z, s.z);
        });
        const float* m = d.cg_dinv + 9 * (size_t)i;
        float3 r = make_float3(fmaf(m[2], s.z, fmaf(m[1], s.y, m[0] * s.x)), fmaf(m[5], s.z, fmaf(m[4], s.y, m[3] * s.x)),
                               fmaf(m[8], s.z, fmaf(m[7], s.y, m[6] * s.x)));
        const float f = c.dt * c.inv_rho0;
        const float4 p = brick_own_load(bk, row, 2);
        d.cg_Ap[i] = make_float4(fmaf(r.x, f, p.x), fmaf(r.y, f, p.y), fmaf(r.z, f, p.z), 0.f);
    });
    brick_finish(d);
}

// |N(i)| per particle / CSR fill (host debug view of for_all_neighbors)
__global__ void __launch_bounds__(SPH_BLOCK) k_neighbor_count(Consts c, Dev d, int* counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    int n = 0;
    for_all_neighbors(c, d, i, d.pv[i], [&](int, float4, float3, float) { n++; });
    counts[i] = n;
}
__global__ void __launch_bounds__(SPH_BLOCK) k_neighbor_fill(Consts c, Dev d, const int* offsets, int* indices) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    int o = offsets[i];
    for_all_neighbors(c, d, i, d.pv[i], [&](int j, float4, float3, float) { indices[o++] = j; });
}

// Launch geometry of a brick kernel: dynamic shared memory = window budget x 16 B x staged arrays, grid = what is
// resident at once (occupancy x SMs) — the CTAs are persistent.  Cached per (device, kernel, bytes).
template <class K>
int brick_grid_for(SphHandle* h, K kernel, size_t smem) {
    static std::map<std::pair<int, std::pair<const void*, size_t>>, int> cache;
    const auto key = std::make_pair((int)h->P.device, std::make_pair((const void*)kernel, smem));
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SPH_BRICK_THREADS, smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->P.device);
    const int grid = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    cache[key] = grid;
    return grid;
}

}  // namespace

#define LAUNCH(...)                                                           \
    do {                                                                      \
        if (h->c.N > 0) {                                                     \
            SphProf _prof(h, #__VA_ARGS__);                                   \
            __VA_ARGS__<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, h->stream>>>(h->c, h->d); \
            h->launches++;                                                    \
        }                                                                     \
    } while (0)

// brick kernel staging NARR arrays; extra kernel arguments follow
#define LAUNCH_BRICK(NARR, KERNEL, ...)                                                          \
    do {                                                                                         \
        if (h->c.N > 0) {                                                                        \
            sph_bricks_refresh(h);                                                               \
            SphProf _prof(h, #KERNEL);                                                           \
            const size_t smem_ = (size_t)h->wmax * 16 * (NARR);                                  \
            const int grid_ = brick_grid_for(h, KERNEL, smem_);                                  \
            KERNEL<<<grid_, SPH_BRICK_THREADS, smem_, h->stream>>>(h->c, h->d, h->wmax, ##__VA_ARGS__); \
            h->launches++;                                                                       \
        }                                                                                        \
    } while (0)

// (re)build the lists without touching densities when a list consumer finds them stale
bool sph_lists_ready(SphHandle* h) {
    if (!h->lists_enabled) return false;
    if (!h->list_valid) {
        LAUNCH_BRICK(1, (kb_build<true, false, false>));
        h->list_valid = true;
    }
    return true;
}

// refresh ghost copies that lag their owners (Z-slabs): contiguous NCCL halo of each stale field
void sph_ghost_sync(SphHandle* h, int what) {
    if (!sph_is_slab(h)) return;
    const int need = h->ghost_stale & what;
    int rc = 0;
    if (!rc && (need & GHOST_PV)) rc = sph_slab_halo(h, h->d.pv, 16);
    if (!rc && (need & GHOST_RHO)) rc = sph_slab_halo(h, h->d.rho, 4);
    // inside a peer loop the writers have signalled their completion: pull the ghosts from the neighbours' memory
    if (!rc && (need & GHOST_VEL)) {
        if (h->peer_signalled & GHOST_VEL) sph_slab_peer_pull(h, GHOST_VEL, true);
        else rc = sph_slab_halo(h, h->d.vm, 16);
    }
    if (!rc && (need & GHOST_AUX)) {
        if (h->peer_signalled & GHOST_AUX) sph_slab_peer_pull(h, GHOST_AUX, true);
        else rc = sph_slab_halo(h, h->d.aux, 16);
    }
    if (rc && !h->sticky_rc) h->sticky_rc = rc;
    h->ghost_stale &= ~need;
    h->peer_signalled &= ~need;
}

static void prep_aux(SphHandle* h, int mode) {
    sph_ghost_dirty(h, GHOST_AUX);
    switch (mode) {
        case AUX_RHO_M: LAUNCH(k_prep_aux<AUX_RHO_M>); break;
        case AUX_KAPPA: LAUNCH(k_prep_aux<AUX_KAPPA>); break;
        case AUX_KAPPA_V: LAUNCH(k_prep_aux<AUX_KAPPA_V>); break;
        case AUX_PRESSURE: LAUNCH(k_prep_aux<AUX_PRESSURE>); break;
    }
}
// (., ., rho, m) of the ghosts is local data once their rho is current: no halo of aux needed
static void prep_aux_rho_m(SphHandle* h) {
    prep_aux(h, AUX_RHO_M);
    h->ghost_stale &= ~GHOST_AUX;
}

void sph_launch_rigid_volume(SphHandle* h) { LAUNCH(k_rigid_volume); sph_ghost_dirty(h, GHOST_PV | GHOST_VEL); }

// compute_density; with_alpha (DFSPH step tail): compute_alpha on the same staged window
void sph_launch_density(SphHandle* h, bool with_alpha) {
    sph_ghost_sync(h, GHOST_PV);
    sph_ghost_dirty(h, GHOST_RHO);
    if (h->lists_enabled) {
        if (with_alpha) LAUNCH_BRICK(1, (kb_build<true, true, true>));
        else LAUNCH_BRICK(1, (kb_build<true, true, false>));
        h->list_valid = true;
    } else {
        if (with_alpha) LAUNCH_BRICK(1, (kb_build<false, true, true>));
        else LAUNCH_BRICK(1, (kb_build<false, true, false>));
    }
}
static void zero_rows(SphHandle* h, float4* a) { cudaMemsetAsync(a, 0, sizeof(float4) * (size_t)h->c.N, h->stream); }
void sph_launch_pressure_accel(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    prep_aux(h, AUX_PRESSURE);
    sph_ghost_sync(h, GHOST_AUX);
    zero_rows(h, h->d.acc);   // the reference zero-fills every row first (base_solver.py:139-141)
    if (sph_lists_ready(h)) LAUNCH_BRICK(2, (kb_pressure_accel<false, true>));
    else LAUNCH_BRICK(2, (kb_pressure_accel<false, false>));
}
void sph_launch_temp_pressure_accel(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    prep_aux(h, AUX_PRESSURE);
    sph_ghost_sync(h, GHOST_AUX);
    zero_rows(h, h->d.a_p);
    if (sph_lists_ready(h)) LAUNCH_BRICK(2, (kb_pressure_accel<true, true>));
    else LAUNCH_BRICK(2, (kb_pressure_accel<true, false>));
}
void sph_launch_surface_tension(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV);   // masses ride in vm.w and never change between sorts
    if (sph_lists_ready(h)) LAUNCH_BRICK(2, (kb_surface_tension<true>));
    else LAUNCH_BRICK(2, (kb_surface_tension<false>));
}
void sph_launch_viscosity(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO | GHOST_VEL);
    prep_aux_rho_m(h);
    if (sph_lists_ready(h)) LAUNCH_BRICK(3, (kb_viscosity<true>));
    else LAUNCH_BRICK(3, (kb_viscosity<false>));
}
void sph_launch_dfsph_alpha(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV);
    if (sph_lists_ready(h)) LAUNCH_BRICK(1, (kb_pv_sweep<true, false>));
    else LAUNCH_BRICK(1, (kb_pv_sweep<false, false>));
}

static SolveCtl solve_ctl(SphHandle* h, int mode, bool speculative, float eta, int fuse_field = 0) {
    SolveCtl ctl;
    ctl.mode = mode;
    ctl.speculative = speculative ? 1 : 0;
    ctl.n_global = h->slab ? (float)h->n_global : (float)h->c.N;
    ctl.eta = eta;
    ctl.peer = sph_slab_peer_links(h, fuse_field);
    return ctl;
}

// Inside a peer loop, when the last writer of `field` signalled its completion: let the sweep read the ghosts of that
// payload straight from the neighbours' memory (boundary bricks come last and wait for the signal there) instead of
// refreshing the local ghost copies first.  The local copies stay marked stale.
static bool fuse_ghost_reads(const SphHandle* h, int field) {
    static const bool enabled = [] { const char* e = getenv("SPH_B200_PEER_FUSED"); return !(e && e[0] == '0'); }();
    return enabled && h->peer_loop && h->lists_enabled && (h->ghost_stale & field) && (h->peer_signalled & field);
}

// fused: + kappa(_v) + aux; mode / speculative / eta: see SolveMode
template <bool STAR>
static void launch_density_change(SphHandle* h, bool fused, int mode, bool speculative, float eta) {
    const bool fuse = fuse_ghost_reads(h, GHOST_VEL);
    sph_ghost_sync(h, GHOST_PV | (STAR ? GHOST_RHO : 0) | (fuse ? 0 : GHOST_VEL));
    if (fused) {
        sph_ghost_dirty(h, GHOST_AUX);
        if (h->peer_loop) h->peer_signalled |= GHOST_AUX; else h->peer_signalled &= ~GHOST_AUX;
    }
    const SolveCtl ctl = solve_ctl(h, fused ? mode : SOLVE_PLAIN, speculative, eta, fuse ? GHOST_VEL : 0);
    if (sph_lists_ready(h)) {
        if (fused) LAUNCH_BRICK(2, (kb_dfsph_density_change<STAR, true, true>), ctl);
        else LAUNCH_BRICK(2, (kb_dfsph_density_change<STAR, true, false>), ctl);
    } else {
        if (fused) LAUNCH_BRICK(2, (kb_dfsph_density_change<STAR, false, true>), ctl);
        else LAUNCH_BRICK(2, (kb_dfsph_density_change<STAR, false, false>), ctl);
    }
}
void sph_launch_dfsph_solve_check(SphHandle* h, float eta) {
    SphProf _prof(h, "k_dfsph_solve_check");
    k_dfsph_solve_check<<<1, 1, 0, h->stream>>>(h->d, solve_ctl(h, SOLVE_TEST, true, eta));
    h->launches++;
}
void sph_launch_dfsph_density_derivative(SphHandle* h, bool fused, int mode, bool speculative, float eta) {
    launch_density_change<false>(h, fused, mode, speculative, eta);
}
void sph_launch_dfsph_density_star(SphHandle* h, bool fused, int mode, bool speculative, float eta) {
    launch_density_change<true>(h, fused, mode, speculative, eta);
}
// aux_mode < 0: the fused density-change kernel has just written aux = (kappa, kappa/rho, rho, m)
static void launch_correct(SphHandle* h, int aux_mode, bool speculative) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    if (aux_mode >= 0) prep_aux(h, aux_mode);
    const bool fuse = fuse_ghost_reads(h, GHOST_AUX);
    if (!fuse) sph_ghost_sync(h, GHOST_AUX);
    const SolveCtl ctl = solve_ctl(h, SOLVE_PLAIN, speculative, 0.0f, fuse ? GHOST_AUX : 0);
    const bool wrench = h->c.has_dynamic_rigid != 0;
    if (sph_lists_ready(h)) { if (wrench) LAUNCH_BRICK(2, (kb_dfsph_correct<true, true>), ctl); else LAUNCH_BRICK(2, (kb_dfsph_correct<true, false>), ctl); }
    else { if (wrench) LAUNCH_BRICK(2, (kb_dfsph_correct<false, true>), ctl); else LAUNCH_BRICK(2, (kb_dfsph_correct<false, false>), ctl); }
    sph_ghost_dirty(h, GHOST_VEL);
    if (h->peer_loop) h->peer_signalled |= GHOST_VEL; else h->peer_signalled &= ~GHOST_VEL;
}
void sph_launch_dfsph_correct_divergence(SphHandle* h, bool aux_ready, bool speculative) { launch_correct(h, aux_ready ? -1 : AUX_KAPPA_V, speculative); }
void sph_launch_dfsph_correct_density(SphHandle* h, bool aux_ready, bool speculative) { launch_correct(h, aux_ready ? -1 : AUX_KAPPA, speculative); }

void sph_launch_pcisph_density_star(SphHandle* h) {
    if (sph_lists_ready(h)) LAUNCH_BRICK(2, (kb_pcisph_density_star<true>));
    else LAUNCH_BRICK(2, (kb_pcisph_density_star<false>));
}
void sph_launch_cg_prepare1(SphHandle* h) {
    prep_aux_rho_m(h);
    if (sph_lists_ready(h)) LAUNCH_BRICK(3, (kb_cg_prepare1<true>));
    else LAUNCH_BRICK(3, (kb_cg_prepare1<false>));
}
void sph_launch_cg_Ap(SphHandle* h, bool aux_ready) {   // aux = (., ., rho, m): unchanged inside the CG loop
    if (!aux_ready) prep_aux_rho_m(h);
    if (sph_lists_ready(h)) LAUNCH_BRICK(3, (kb_cg_Ap<true>));
    else LAUNCH_BRICK(3, (kb_cg_Ap<false>));
}

void sph_launch_neighbor_count(SphHandle* h, int* counts) {
    if (h->c.N <= 0) return;
    SphProf _prof(h, "k_neighbor_count");
    k_neighbor_count<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, h->stream>>>(h->c, h->d, counts);
    h->launches++;
}
void sph_launch_neighbor_fill(SphHandle* h, const int* offsets, int* indices) {
    if (h->c.N <= 0) return;
    k_neighbor_fill<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, h->stream>>>(h->c, h->d, offsets, indices);
    h->launches++;
}
