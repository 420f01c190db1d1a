// sph_sweeps.cu — per-particle neighbour summations ("sweeps"), one thread per particle.
//
// Positions are frozen between a sort and the next position update and every solver runs many sweeps
// in that interval (DFSPH: density, alpha, the divergence solve, then next step's surface tension,
// viscosity and the ~25-iteration density solve).  So:
//   * compute_density, the first sweep after every sort, tests the ~250 candidates of each fluid
//     particle's 27-cell window once — candidates staged in shared memory by TMA bulk copies
//     (sph_window.cuh) — and records the ~40 accepted neighbours in an ELL list nbr[k][i];
//   * every later sweep streams its list (coalesced along i) and fetches all it needs about
//     neighbour j with ONE 256-bit gather of a 32-byte record {pv_j, payload_j} (LDG.E.256, sm_100+):
//     the sweeps are bound by L1 wavefronts (one per distinct 128-byte line per gather), so halving
//     or thirding the gathers per pair is what makes them fast (profiles/r01_*);
//   * rows that overflow the list, and everything when lists are disabled (SPH_B200_NO_LISTS=1),
//     re-derive their neighbours by walking the window in global memory (same order, same result).
//
// Each kernel replaces one @ti.kernel + its *_task of the reference (cited per kernel); the task
// bodies are inlined lambdas where upstream passes ti.template() callbacks into for_all_neighbors
// (base_container.py:549-560).
#include <map>

#include "sph_kernels.h"
#include "sph_window.cuh"

namespace {

extern __shared__ __align__(16) float4 dyn_smem[];

// Compile-time tuning knobs for variant builds (`make variants`, tools/tune_variants.sh).  Undefined, the
// preprocessed source -- and therefore the shipped binary -- is exactly what was validated on hardware.
//   SPH_CORRECT_MINB   min resident CTAs per SM for k_dfsph_correct (caps its registers: 56 -> 48 / 40)
//   SPH_LIST_UNROLL8   eight index -> record gather chains in flight in rec_neighbors instead of four
//   SPH_IDX_NO_ALLOCATE / SPH_REC_EVICT_LAST   L1 allocation hints for the list stream / the record gathers (sph_common.cuh)
//   SPH_DEVICE_CONVERGENCE   loop-exit test of the DFSPH density solve on the device, host reads once per batch (sph_api.cu)
#ifdef SPH_CORRECT_MINB
#define SPH_CORRECT_BOUNDS __launch_bounds__(SPH_BLOCK, SPH_CORRECT_MINB)
#else
#define SPH_CORRECT_BOUNDS __launch_bounds__(SPH_BLOCK)
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

// All neighbours j of particle i in walk order; visit(j, pj, hj, R, r2) with hj = rec[j].hi.
// LIST: stream the list recorded by the density pass (4 index -> record load chains in flight).
template <bool LIST, class Visit>
__device__ __forceinline__ void rec_neighbors(const Consts& c, const Dev& d, const Rec* __restrict__ rec, int i, float4 pi,
                                              Visit&& visit) {
    if (LIST) {
        const int n = d.nbr_count[i];
        if (n <= d.nbr_kmax) {
            const int* __restrict__ col = d.nbr + i;
            const size_t stride = (size_t)d.nbr_stride;
            int k = 0;
#ifdef SPH_LIST_UNROLL8
            for (; k + 8 <= n; k += 8) {
                int j[8];
                float4 p[8], hh[8];
#pragma unroll
                for (int u = 0; u < 8; u++) j[u] = SPH_LDG_IDX(col + (size_t)(k + u) * stride);
#pragma unroll
                for (int u = 0; u < 8; u++) ldg_rec(rec + j[u], p[u], hh[u]);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float3 R = make_float3(pi.x - p[u].x, pi.y - p[u].y, pi.z - p[u].z);
                    visit(j[u], p[u], hh[u], R, dist2(R));
                }
            }
#endif
            for (; k + 4 <= n; k += 4) {
                const int j0 = SPH_LDG_IDX(col + (size_t)k * stride), j1 = SPH_LDG_IDX(col + (size_t)(k + 1) * stride);
                const int j2 = SPH_LDG_IDX(col + (size_t)(k + 2) * stride), j3 = SPH_LDG_IDX(col + (size_t)(k + 3) * stride);
                float4 p0, h0, p1, h1, p2, h2, p3, h3;
                ldg_rec(rec + j0, p0, h0); ldg_rec(rec + j1, p1, h1); ldg_rec(rec + j2, p2, h2); ldg_rec(rec + j3, p3, h3);
                float3 R;
                R = make_float3(pi.x - p0.x, pi.y - p0.y, pi.z - p0.z); visit(j0, p0, h0, R, dist2(R));
                R = make_float3(pi.x - p1.x, pi.y - p1.y, pi.z - p1.z); visit(j1, p1, h1, R, dist2(R));
                R = make_float3(pi.x - p2.x, pi.y - p2.y, pi.z - p2.z); visit(j2, p2, h2, R, dist2(R));
                R = make_float3(pi.x - p3.x, pi.y - p3.y, pi.z - p3.z); visit(j3, p3, h3, R, dist2(R));
            }
            for (; k < n; k++) {
                const int j = SPH_LDG_IDX(col + (size_t)k * stride);
                float4 pj, hj;
                ldg_rec(rec + j, pj, hj);
                const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                visit(j, pj, hj, R, dist2(R));
            }
            return;
        }
    }
    for_all_neighbors(c, d, i, pi, [&](int j, float4 pj, float3 R, float r2) { visit(j, pj, __ldg(&rec[j].hi), R, r2); });
}

// position-only variant (alpha, PCISPH density): 128-bit gathers of pv
template <bool LIST, class Visit>
__device__ __forceinline__ void pv_neighbors(const Consts& c, const Dev& d, int i, float4 pi, Visit&& visit) {
    if (LIST) {
        const int n = d.nbr_count[i];
        if (n <= d.nbr_kmax) {
            const int* __restrict__ col = d.nbr + i;
            const size_t stride = (size_t)d.nbr_stride;
            int k = 0;
            for (; k + 4 <= n; k += 4) {
                const int j0 = __ldg(col + (size_t)k * stride), j1 = __ldg(col + (size_t)(k + 1) * stride);
                const int j2 = __ldg(col + (size_t)(k + 2) * stride), j3 = __ldg(col + (size_t)(k + 3) * stride);
                const float4 p0 = __ldg(d.pv + j0), p1 = __ldg(d.pv + j1), p2 = __ldg(d.pv + j2), p3 = __ldg(d.pv + j3);
                float3 R;
                R = make_float3(pi.x - p0.x, pi.y - p0.y, pi.z - p0.z); visit(j0, p0, R, dist2(R));
                R = make_float3(pi.x - p1.x, pi.y - p1.y, pi.z - p1.z); visit(j1, p1, R, dist2(R));
                R = make_float3(pi.x - p2.x, pi.y - p2.y, pi.z - p2.z); visit(j2, p2, R, dist2(R));
                R = make_float3(pi.x - p3.x, pi.y - p3.y, pi.z - p3.z); visit(j3, p3, R, dist2(R));
            }
            for (; k < n; k++) {
                const int j = __ldg(col + (size_t)k * stride);
                const float4 pj = __ldg(d.pv + j);
                const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                visit(j, pj, R, dist2(R));
            }
            return;
        }
    }
    for_all_neighbors(c, d, i, pi, visit);
}

// recB[i].hi = (s0, s1, rho_i, m_i), the scalar payload of the pressure / correction / tension sweeps
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
    d.recB[i].hi = make_float4(s0, s1, rho, m);
}

// refresh the record copies of pv / vm (host-side edits, kernels that do not write the records)
__global__ void __launch_bounds__(SPH_BLOCK) k_sync_records(Consts c, Dev d, int with_vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    const float4 p = d.pv[i];
    d.recA[i].lo = p;
    d.recB[i].lo = p;
    // ghost velocities live in recA only (refreshed by halos), never re-derived from the ghosts' vm
    if (with_vel && SPH_IS_ROW(c, i)) d.recA[i].hi = d.vm[i];
}

// compute_rigid_particle_volume (base_solver.py:105-123); rigid rows, plain window walk in global
// memory (once per step over the boundary shell)
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
    reinterpret_cast<float*>(&d.recA[i].lo)[3] = -V;
    reinterpret_cast<float*>(&d.recB[i].lo)[3] = -V;
    reinterpret_cast<float*>(&d.recA[i].hi)[3] = m;
}

// compute_density (base_solver.py:521-541) fused with the neighbour-list build; candidates come
// from the TMA-staged window.  DENSITY: write rho; BUILD: record the accepted neighbours.
template <bool DENSITY, bool BUILD>
__global__ void __launch_bounds__(SPH_BLOCK) k_density(Consts c, Dev d, int wmax) {
    __shared__ int s_desc[SPH_DESC_INTS];
    __shared__ unsigned long long s_mbar;
    const Window win = window_open(d, dyn_smem, wmax, s_desc, &s_mbar);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) {
        if (BUILD) d.nbr_count[i] = 0;
        return;
    }
    float ret = 0.0f;
    int n = 0;
    int* col = d.nbr + i;
    const size_t stride = (size_t)d.nbr_stride;
    const int kmax = d.nbr_kmax;
    auto body = [&](int j, float4 pj, float3, float r2) {
        if (DENSITY) ret += fabsf(pj.w) * kernel_W_q(c, sqrtf(r2) * c.inv_h);
        if (BUILD) {
            if (n < kmax) col[(size_t)n * stride] = j;
            n++;
        }
    };
    if (win.staged) window_walk(c, d, win, i, pi, body);
    else for_all_neighbors(c, d, i, pi, body);
    if (DENSITY) d.rho[i] = (pi.w * c.kW + ret) * c.rho0;
    if (BUILD) d.nbr_count[i] = n;   // may exceed kmax: such rows re-derive their neighbours
}

// compute_pressure_acceleration (base_solver.py:135-187) and, with TEMP, PCISPH's
// compute_temp_pressure_acceleration (PCISPH.py:74-107: fluid rows, no rigid wrench, output a_p).
// recB.hi = (p / rho^2, p, rho, m)
template <bool TEMP, bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_pressure_accel(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    float4* out = TEMP ? d.a_p : d.acc;
    bool active = pi.w > 0.0f;
    if (!TEMP) active = active && d.is_dynamic[i] != 0;
    if (!active) {
        out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float dpi = d.recB[i].hi.x;
    float3 ret = make_float3(0.f, 0.f, 0.f);
    rec_neighbors<LIST>(c, d, d.recB, i, pi, [&](int j, float4 pj, float4 aj, float3 R, float r2) {
        const float gs = kernel_gradient_scale(c, r2);
        float coef;
        if (pj.w > 0.0f) {
            coef = -aj.w * (dpi + aj.x);
        } else {
            coef = -c.rho0 * (-pj.w) * dpi;
            if (!TEMP && c.has_dynamic_rigid && __ldg(d.is_dynamic + j)) {
                float3 force = R * (-coef * gs * (c.rho0 * pi.w));
                add_wrench(d, __ldg(d.object_id + j), force, f3(pi));  // arm from x_i (base_solver.py:185)
            }
        }
        const float s = coef * gs;
        ret.x = fmaf(s, R.x, ret.x); ret.y = fmaf(s, R.y, ret.y); ret.z = fmaf(s, R.z, ret.z);
    });
    out[i] = make_float4(ret.x, ret.y, ret.z, 0.f);
}

// compute_surface_tension_acceleration (base_solver.py:209-229): a_i += sum.  recB.hi.w = m_j
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_surface_tension(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float sm = c.sigma / d.recB[i].hi.w;
    float3 a = make_float3(0.f, 0.f, 0.f);
    rec_neighbors<LIST>(c, d, d.recB, i, pi, [&](int, float4 pj, float4 aj, float3 R, float r2) {
        if (!(pj.w > 0.0f)) return;
        const float w = r2 > c.diameter2 ? kernel_W_q(c, sqrtf(r2) * c.inv_h) : c.w_diameter;
        const float s = sm * aj.w * w;
        a.x = fmaf(-s, R.x, a.x); a.y = fmaf(-s, R.y, a.y); a.z = fmaf(-s, R.z, a.z);
    });
    float4 acc = d.acc[i];
    d.acc[i] = make_float4(acc.x + a.x, acc.y + a.y, acc.z + a.z, 0.f);
}

// compute_viscosity_acceleration_standard (base_solver.py:231-278): a_i += sum / rho0.
// recA.hi = (v_j, m_j); rho_j of fluid neighbours is one more scalar gather
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_viscosity(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 vi = d.vm[i];
    const float den_i = d.rho[i];
    float3 a = make_float3(0.f, 0.f, 0.f);
    rec_neighbors<LIST>(c, d, d.recA, i, pi, [&](int j, float4 pj, float4 vj, float3 R, float r2) {
        const float v_xy = dot3(make_float3(vi.x - vj.x, vi.y - vj.y, vi.z - vj.z), R);
        const float gs = kernel_gradient_scale(c, r2);
        float coef;
        if (pj.w > 0.0f) {
            coef = c.visc_cf * ((vi.w + vj.w) * 0.5f) / __ldg(d.rho + j);
        } else {
            coef = c.visc_cb * (c.rho0 * (-pj.w)) / den_i;
        }
        const float s = coef / (r2 + c.visc_eps) * v_xy * gs;
        a.x = fmaf(s, R.x, a.x); a.y = fmaf(s, R.y, a.y); a.z = fmaf(s, R.z, a.z);
        if (!(pj.w > 0.0f) && c.has_dynamic_rigid && __ldg(d.is_dynamic + j)) {
            float3 force = R * (-s * vi.w * c.inv_rho0);     // -acc * m_i / rho0
            add_wrench(d, __ldg(d.object_id + j), force, f3(pj));
        }
    });
    float4 acc = d.acc[i];
    d.acc[i] = make_float4(acc.x + a.x * c.inv_rho0, acc.y + a.y * c.inv_rho0, acc.z + a.z * c.inv_rho0, 0.f);
}

// DFSPH compute_alpha (DFSPH.py:22-62)
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_alpha(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    float3 grad_i = make_float3(0.f, 0.f, 0.f);
    float sum_k = 0.0f;
    pv_neighbors<LIST>(c, d, i, pi, [&](int, float4 pj, float3 R, float r2) {
        const float s = -fabsf(pj.w) * kernel_gradient_scale(c, r2);
        const float3 g = R * s;
        if (pj.w > 0.0f) sum_k += dist2(g);
        grad_i = grad_i + g;
    });
    sum_k += dist2(grad_i);
    d.alpha[i] = sum_k > 1e-5f ? 1.0f / sum_k : 0.0f;
}

// DFSPH compute_density_derivative (DFSPH.py:65-101) / compute_density_star (:104-126).
// FUSED (the library's own solver loops): also the kappa of the next correction step
// (compute_kappa_v :132-137 / compute_kappa :217-223, written to the field and to recB.hi) and the
// error sum (compute_density_derivative_error :205-211 / compute_density_error :285-294).
template <bool STAR, bool LIST, bool FUSED>
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_density_change(Consts c, Dev d) {
#ifdef SPH_DEVICE_CONVERGENCE
    if (d.red[CTRL_DONE] != 0.0) return;   // uniform over the grid: no thread reaches the block reduction
#endif
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float err = 0.0f;
    if (SPH_IS_ROW(c, i)) {
        const float4 pi = d.pv[i];
        if (pi.w > 0.0f) {
            const float4 vi = d.vm[i];
            float delta = 0.0f;
            int nn = 0;
            rec_neighbors<LIST>(c, d, d.recA, i, pi, [&](int, float4 pj, float4 vj, float3 R, float r2) {
                const float vr = dot3(make_float3(vi.x - vj.x, vi.y - vj.y, vi.z - vj.z), R);
                delta = fmaf(fabsf(pj.w) * kernel_gradient_scale(c, r2), vr, delta);
                nn++;
            });
            const float rho = d.rho[i];
            float kap;
            if (STAR) {
                const float rs = fmaxf(rho / c.rho0 + c.dt * delta, 1.0f);
                d.rho_star[i] = rs;
                kap = (rs - 1.0f) * d.alpha[i] * c.inv_dt;
                if (FUSED) { d.kappa[i] = kap; err = rs - 1.0f; }
            } else {
                float adv = fmaxf(delta, 0.0f);
                if (nn < 20) adv = 0.0f;   // particle deficiency (DFSPH.py:93-95)
                d.drho[i] = adv;
                kap = adv * d.alpha[i];
                if (FUSED) { d.kappa_v[i] = kap; err = c.rho0 * adv; }
            }
            if (FUSED) d.recB[i].hi = make_float4(kap, kap / rho, rho, vi.w);
        }
    }
    if (FUSED) block_reduce_add(d.red + RED_ERR, (double)err);
}

// DFSPH correct_divergence_step (DFSPH.py:161-202) / correct_density_error_step (:245-283).
// recB.hi = (kappa_j, kappa_j / rho_j, rho_j, m_j); the new velocity goes to vm and to recA.hi
template <bool LIST>
__global__ void SPH_CORRECT_BOUNDS k_dfsph_correct(Consts c, Dev d) {
#ifdef SPH_DEVICE_CONVERGENCE
    if (d.red[CTRL_DONE] != 0.0) return;   // the solve converged earlier in this batch: a speculative launch does nothing
#endif
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 ai = d.recB[i].hi;
    const float k_i = ai.x, ki_rho = ai.y;
    const float thresh = 1e-5f * c.dt;   // m_eps * dt
    const bool rigid_on = fabsf(k_i) > thresh;
    float3 dv = make_float3(0.f, 0.f, 0.f);
    rec_neighbors<LIST>(c, d, d.recB, i, pi, [&](int j, float4 pj, float4 aj, float3 R, float r2) {
        float s;
        if (pj.w > 0.0f) {
            if (!(fabsf(k_i + aj.x) > thresh)) return;
            s = pj.w * kernel_gradient_scale(c, r2) * (ki_rho + aj.y) * c.rho0;
        } else {
            if (!rigid_on) return;
            s = (-pj.w) * kernel_gradient_scale(c, r2) * ki_rho * c.rho0;
            if (c.has_dynamic_rigid && __ldg(d.is_dynamic + j)) {
                float3 force = R * (s * c.inv_dt * (pi.w * c.rho0));
                add_wrench(d, __ldg(d.object_id + j), force, f3(pj));
            }
        }
        dv.x = fmaf(-s, R.x, dv.x); dv.y = fmaf(-s, R.y, dv.y); dv.z = fmaf(-s, R.z, dv.z);
    });
    float4 v = d.vm[i];
    v = make_float4(v.x + dv.x, v.y + dv.y, v.z + dv.z, v.w);
    d.vm[i] = v;
    d.recA[i].hi = v;
}

#ifdef SPH_DEVICE_CONVERGENCE
// The loop-exit test of DFSPH.correct_density_error (DFSPH.py:236-242) on the device, with the host's arithmetic
// (f32 division of the f64 sum by the particle count, compare with eta); also clears the error slot for the next
// iteration (the host path does that with a memset).
__global__ void k_dfsph_solve_check(Dev d, float n_global, float eta) {
    if (d.red[CTRL_DONE] == 0.0) {
        const float e = (float)d.red[RED_ERR] / n_global;
        d.red[CTRL_ITERS] += 1.0;
        d.red[CTRL_ERR] = (double)e;
        if (e <= eta) d.red[CTRL_DONE] = 1.0;
    }
    d.red[RED_ERR] = 0.0;
}
#endif

// PCISPH compute_density_star (PCISPH.py:32-62): predicted positions, no self term, neighbour
// set from the current positions.  Accumulates sum max(0, rho*/rho0 - 1) into red[RED_ERR].
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_pcisph_density_star(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float err = 0.0f;
    if (SPH_IS_ROW(c, i)) {
        const float4 pi = d.pv[i];
        if (pi.w > 0.0f) {
            const float4 xi = d.x_pred[i];
            float ret = 0.0f;
            pv_neighbors<LIST>(c, d, i, pi, [&](int j, float4 pj, float3, float) {
                float4 xj = pj.w > 0.0f ? __ldg(d.x_pred + j) : pj;
                float r2 = dist2(make_float3(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z));
                ret += fabsf(pj.w) * kernel_W_q(c, sqrtf(r2) * c.inv_h);
            });
            d.rho_star[i] = ret * c.rho0;
            err = fmaxf(0.0f, ret - 1.0f);
        }
    }
    block_reduce_add(d.red + RED_ERR, (double)err);
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
// D_i^-1, b_i and p_i <- x_i.  recA.hi = (v_j, m_j)
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_prepare1(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 vi = d.vm[i];
    const float den_i = d.rho[i];
    // ret = -sum A_ij (symmetric in R (x) R): 6 unique entries
    float sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    float3 b = make_float3(0.f, 0.f, 0.f);
    rec_neighbors<LIST>(c, d, d.recA, i, pi, [&](int j, float4 pj, float4 vj, float3 R, float r2) {
        const float rho_j = pj.w > 0.0f ? __ldg(d.rho + j) : 1.0f;
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
    const float m00 = 1.0f - sxx * f, m01 = -sxy * f, m02 = -sxz * f, m11 = 1.0f - syy * f, m12 = -syz * f, m22 = 1.0f - szz * f;
    const float c00 = m11 * m22 - m12 * m12, c01 = m12 * m02 - m01 * m22, c02 = m01 * m12 - m11 * m02;
    const float inv = 1.0f / (m00 * c00 + m01 * c01 + m02 * c02);
    float* o = d.cg_dinv + 9 * (size_t)i;
    o[0] = c00 * inv; o[1] = c01 * inv; o[2] = c02 * inv;
    o[3] = c01 * inv; o[4] = (m00 * m22 - m02 * m02) * inv; o[5] = (m02 * m01 - m00 * m12) * inv;
    o[6] = c02 * inv; o[7] = o[5]; o[8] = (m00 * m11 - m01 * m01) * inv;
    d.cg_b[i] = make_float4(vi.x - c.dt * b.x * c.inv_rho0, vi.y - c.dt * b.y * c.inv_rho0, vi.z - c.dt * b.z * c.inv_rho0, 0.f);
    d.cg_p[i] = d.cg_x[i];
}

// compute_Ap (base_solver.py:373-391): Ap_i = p_i + dt/rho0 * D_i^-1 sum_{fluid j} (-A_ij) p_j.
// recB.hi = (., ., rho_j, m_j); cg_p[j] is a second 128-bit gather
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_Ap(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    SPH_ROW_OR_RETURN(c, i);
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 ai = d.recB[i].hi;
    const float mi = ai.w, den_i = ai.z;
    float3 s = make_float3(0.f, 0.f, 0.f);
    rec_neighbors<LIST>(c, d, d.recB, i, pi, [&](int j, float4 pj, float4 aj, float3 R, float r2) {
        if (!(pj.w > 0.0f)) return;
        const float a = -visc_A_scale(c, mi, den_i, pj, aj.w, aj.z, r2);
        const float4 pj_cg = __ldg(d.cg_p + j);
        const float t = a * dot3(R, f3(pj_cg));   // (-A_ij) p_j = a R (R . p_j)
        s.x = fmaf(t, R.x, s.x); s.y = fmaf(t, R.y, s.y); s.z = fmaf(t, R.z, s.z);
    });
    const float* m = d.cg_dinv + 9 * (size_t)i;
    float3 r = make_float3(m[0] * s.x + m[1] * s.y + m[2] * s.z, m[3] * s.x + m[4] * s.y + m[5] * s.z,
                           m[6] * s.x + m[7] * s.y + m[8] * s.z);
    const float f = c.dt * c.inv_rho0;
    const float4 p = d.cg_p[i];
    d.cg_Ap[i] = make_float4(fmaf(r.x, f, p.x), fmaf(r.y, f, p.y), fmaf(r.z, f, p.z), 0.f);
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

template <class K>
void set_smem_limit(K kernel, size_t bytes) {
    static std::map<const void*, size_t> current;   // per kernel entry point
    size_t& cur = current[(const void*)kernel];
    if (bytes > cur) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        cur = bytes;
    }
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

// list-build kernel: one CTA per chunk, dynamic shared memory = wmax window slots x 16 B
#define LAUNCH_WIN(...)                                                       \
    do {                                                                      \
        if (h->c.N > 0) {                                                     \
            SphProf _prof(h, #__VA_ARGS__);                                   \
            const size_t smem_ = (size_t)h->wmax * 16;                        \
            set_smem_limit(__VA_ARGS__, smem_);                               \
            __VA_ARGS__<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, smem_, h->stream>>>(h->c, h->d, h->wmax); \
            h->launches++;                                                    \
        }                                                                     \
    } while (0)

#define LAUNCH_LIST(kernel, ...)                                              \
    do {                                                                      \
        if (sph_lists_ready(h)) LAUNCH(kernel<__VA_ARGS__ true>);             \
        else LAUNCH(kernel<__VA_ARGS__ false>);                               \
    } while (0)

// (re)build the lists without touching densities when a list consumer finds them stale
bool sph_lists_ready(SphHandle* h) {
    if (!h->lists_enabled) return false;
    if (!h->list_valid) {
        LAUNCH_WIN(k_density<false, true>);
        h->list_valid = true;
    }
    return true;
}

// record copies of pv (and vm) current?
static void ensure_records(SphHandle* h, bool need_vel) {
    if (h->rec_pos_valid && (!need_vel || h->rec_vel_valid)) return;
    if (h->c.N > 0) {
        SphProf _prof(h, "k_sync_records");
        const int with_vel = !h->rec_vel_valid;
        k_sync_records<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, h->stream>>>(h->c, h->d, with_vel);
        h->launches++;
        if (with_vel) sph_ghost_dirty(h, GHOST_VEL);
    }
    h->rec_pos_valid = true;
    h->rec_vel_valid = true;
}

// refresh ghost copies that lag their owners (Z-slabs): contiguous NCCL halo of each stale field
void sph_ghost_sync(SphHandle* h, int what) {
    if (!sph_is_slab(h)) return;
    const int need = h->ghost_stale & what;
    int rc = 0;
    if (!rc && (need & GHOST_PV)) {
        rc = sph_slab_halo(h, h->d.pv, 16);
        h->rec_pos_valid = false;   // record copies of the ghosts' pv are refreshed by ensure_records
    }
    if (!rc && (need & GHOST_RHO)) rc = sph_slab_halo(h, h->d.rho, 4);
    if (!rc && (need & GHOST_VEL)) rc = sph_slab_halo(h, h->d.recA, 32);
    if (!rc && (need & GHOST_AUX)) rc = sph_slab_halo(h, h->d.recB, 32);
    if (rc && !h->sticky_rc) h->sticky_rc = rc;
    h->ghost_stale &= ~need;
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

void sph_launch_rigid_volume(SphHandle* h) { LAUNCH(k_rigid_volume); sph_ghost_dirty(h, GHOST_PV); }
void sph_launch_density(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV);
    sph_ghost_dirty(h, GHOST_RHO);
    if (h->lists_enabled) {
        LAUNCH_WIN(k_density<true, true>);
        h->list_valid = true;
    } else {
        LAUNCH_WIN(k_density<true, false>);
    }
}
void sph_launch_pressure_accel(SphHandle* h) { sph_ghost_sync(h, GHOST_PV); ensure_records(h, false); prep_aux(h, AUX_PRESSURE); sph_ghost_sync(h, GHOST_AUX); LAUNCH_LIST(k_pressure_accel, false, ); }
void sph_launch_temp_pressure_accel(SphHandle* h) { sph_ghost_sync(h, GHOST_PV); ensure_records(h, false); prep_aux(h, AUX_PRESSURE); sph_ghost_sync(h, GHOST_AUX); LAUNCH_LIST(k_pressure_accel, true, ); }
void sph_launch_surface_tension(SphHandle* h) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    ensure_records(h, false);
    prep_aux(h, AUX_RHO_M);
    h->ghost_stale &= ~GHOST_AUX;   // the ghosts' (rho, m) are local data: nothing to fetch
    LAUNCH_LIST(k_surface_tension, );
}
void sph_launch_viscosity(SphHandle* h, bool) { sph_ghost_sync(h, GHOST_PV | GHOST_RHO); ensure_records(h, true); sph_ghost_sync(h, GHOST_VEL); LAUNCH_LIST(k_viscosity, ); }
void sph_launch_dfsph_alpha(SphHandle* h) { sph_ghost_sync(h, GHOST_PV); LAUNCH_LIST(k_dfsph_alpha, ); }
void sph_launch_dfsph_density_derivative(SphHandle* h, bool fused) {
    sph_ghost_sync(h, GHOST_PV);
    ensure_records(h, true);
    sph_ghost_sync(h, GHOST_VEL);
    if (fused) sph_ghost_dirty(h, GHOST_AUX);
    if (sph_lists_ready(h)) { if (fused) LAUNCH(k_dfsph_density_change<false, true, true>); else LAUNCH(k_dfsph_density_change<false, true, false>); }
    else { if (fused) LAUNCH(k_dfsph_density_change<false, false, true>); else LAUNCH(k_dfsph_density_change<false, false, false>); }
}
void sph_launch_dfsph_density_star(SphHandle* h, bool fused) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    ensure_records(h, true);
    sph_ghost_sync(h, GHOST_VEL);
    if (fused) sph_ghost_dirty(h, GHOST_AUX);
    if (sph_lists_ready(h)) { if (fused) LAUNCH(k_dfsph_density_change<true, true, true>); else LAUNCH(k_dfsph_density_change<true, true, false>); }
    else { if (fused) LAUNCH(k_dfsph_density_change<true, false, true>); else LAUNCH(k_dfsph_density_change<true, false, false>); }
}
// aux_ready: the fused density-change kernel has just written recB.hi = (kappa, kappa/rho, rho, m)
void sph_launch_dfsph_correct_divergence(SphHandle* h, bool aux_ready) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    ensure_records(h, false);
    if (!aux_ready) prep_aux(h, AUX_KAPPA_V);
    sph_ghost_sync(h, GHOST_AUX);
    LAUNCH_LIST(k_dfsph_correct, );
    sph_ghost_dirty(h, GHOST_VEL);
}
void sph_launch_dfsph_correct_density(SphHandle* h, bool aux_ready) {
    sph_ghost_sync(h, GHOST_PV | GHOST_RHO);
    ensure_records(h, false);
    if (!aux_ready) prep_aux(h, AUX_KAPPA);
    sph_ghost_sync(h, GHOST_AUX);
    LAUNCH_LIST(k_dfsph_correct, );
    sph_ghost_dirty(h, GHOST_VEL);
}
#ifdef SPH_DEVICE_CONVERGENCE
void sph_launch_dfsph_solve_check(SphHandle* h, float n_global, float eta) {
    SphProf _prof(h, "k_dfsph_solve_check");
    k_dfsph_solve_check<<<1, 1, 0, h->stream>>>(h->d, n_global, eta);
    h->launches++;
}
#endif
void sph_launch_pcisph_density_star(SphHandle* h) { LAUNCH_LIST(k_pcisph_density_star, ); }
void sph_launch_cg_prepare1(SphHandle* h) { ensure_records(h, true); LAUNCH_LIST(k_cg_prepare1, ); }
void sph_launch_cg_Ap(SphHandle* h, bool aux_ready) {   // recB.hi = (., ., rho, m): unchanged inside the CG loop
    ensure_records(h, false);
    if (!aux_ready) prep_aux(h, AUX_RHO_M);
    LAUNCH_LIST(k_cg_Ap, );
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
