// sph_sweeps.cu — per-particle neighbour summations ("sweeps").
//
// Design (see sph_window.cuh): one CTA per chunk of 128 consecutive sorted particles; the chunk's
// 27-cell neighbourhood (<= 9 contiguous ranges of the sorted SoA) is staged into shared memory by
// TMA bulk copies of every float4 payload array the task reads (pv always; vm / aux / x_pred / cg_p
// per task); one thread per particle then accumulates its sum from shared memory.
//
// Positions are frozen between a sort and the next position update and every solver runs many sweeps
// in that interval (DFSPH: density, alpha, the divergence solve, then next step's surface tension,
// viscosity and the ~25-iteration density solve).  The first sweep after a sort (compute_density)
// tests the ~250 window candidates of every fluid particle once and records the ~40 accepted ones as
// 16-bit window slots in an ELL list (nbr16[k][i], coalesced along i); every later sweep streams
// that list.  Rows that overflow the list and chunks whose window exceeds the shared-memory budget
// re-derive their neighbours (same order, same result).
//
// Each kernel replaces one @ti.kernel + its *_task of the reference (cited per kernel); the task
// bodies are generic lambdas where upstream passes ti.template() callbacks into for_all_neighbors
// (base_container.py:549-560).
#include <map>

#include "sph_kernels.h"
#include "sph_window.cuh"

namespace {

extern __shared__ __align__(16) float4 dyn_smem[];

#define WINDOW_PROLOGUE(NPAY, ...)                                                   \
    __shared__ int s_desc[SPH_DESC_INTS];                                            \
    __shared__ unsigned long long s_mbar;                                            \
    const float4* const payload_[NPAY] = {__VA_ARGS__};                              \
    const Window<NPAY> win = window_open<NPAY>(d, payload_, dyn_smem, wmax, s_desc, &s_mbar); \
    const int i = blockIdx.x * blockDim.x + threadIdx.x

template <int NPAY>
__device__ __forceinline__ int to_global(const Window<NPAY>& win, const SmemView<NPAY>&, int idx) { return win.global_index(idx); }
template <int NPAY>
__device__ __forceinline__ int to_global(const Window<NPAY>&, const GlobalView<NPAY>&, int idx) { return idx; }

__device__ __forceinline__ float mass_of(const Dev& d, int j) { return __ldg(reinterpret_cast<const float*>(d.vm + j) + 3); }

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

// aux[i] = (s0, s1, rho_i, m_i), the per-sweep scalar payload staged next to pv
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
// memory (runs once per step over the boundary shell; skipped entirely for static scenes)
__global__ void __launch_bounds__(SPH_BLOCK) k_rigid_volume(Consts c, Dev d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    float4 pi = d.pv[i];
    if (!(pi.w < 0.0f) || !(pi.y <= c.g_upper)) return;
    const int obj_i = d.object_id[i];
    float ret = c.kW;  // W(0)
    for_all_neighbors(c, d, i, pi, [&](int j, float4, float3, float r2) {
        if (__ldg(d.object_id + j) == obj_i) ret += kernel_W_q(c, sqrtf(r2) * c.inv_h);
    });
    float V = 1.0f / ret;
    reinterpret_cast<float*>(d.pv + i)[3] = -V;
    reinterpret_cast<float*>(d.vm + i)[3] = c.rho0 * V;
}

// compute_density (base_solver.py:521-541) fused with the neighbour-list build.
// DENSITY: write rho; BUILD: record the accepted neighbours (walk order) and their count.
template <bool DENSITY, bool BUILD>
__global__ void __launch_bounds__(SPH_BLOCK) k_density(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(1, d.pv);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) {
        if (BUILD) d.nbr_count[i] = 0;
        return;
    }
    float ret = 0.0f;
    int n = 0;
    unsigned short* col = d.nbr16 + i;
    const size_t stride = (size_t)d.nbr_stride;
    const int kmax = d.nbr_kmax;
    if (win.staged) {
        window_walk(c, d, win, i, pi, [&](const SmemView<1>&, int w, float4 pj, float3, float r2) {
            if (DENSITY) ret += fabsf(pj.w) * kernel_W_q(c, sqrtf(r2) * c.inv_h);
            if (BUILD) {
                if (n < kmax) col[(size_t)n * stride] = (unsigned short)w;
                n++;
            }
        });
    } else {
        for_all_neighbors(c, d, i, pi, [&](int, float4 pj, float3, float r2) {
            if (DENSITY) ret += fabsf(pj.w) * kernel_W_q(c, sqrtf(r2) * c.inv_h);
        });
        n = 0x7fffffff;   // no list for this chunk: consumers walk the window
    }
    if (DENSITY) d.rho[i] = (pi.w * c.kW + ret) * c.rho0;
    if (BUILD) d.nbr_count[i] = n;   // may exceed kmax: such rows re-derive their neighbours
}

// compute_pressure_acceleration (base_solver.py:135-187) and, with TEMP, PCISPH's
// compute_temp_pressure_acceleration (PCISPH.py:74-107: fluid rows, no rigid wrench, output a_p).
// aux = (p / rho^2, p, rho, m)
template <bool TEMP, bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_pressure_accel(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(2, d.pv, d.aux);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    float4* out = TEMP ? d.a_p : d.acc;
    bool active = pi.w > 0.0f;
    if (!TEMP) active = active && d.is_dynamic[i] != 0;
    if (!active) {
        out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float dpi = d.aux[i].x;
    float3 ret = make_float3(0.f, 0.f, 0.f);
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
        const float gs = kernel_gradient_scale(c, r2);
        float coef;
        if (pj.w > 0.0f) {
            const float4 aj = view.get(1, idx);
            coef = -aj.w * (dpi + aj.x);
        } else {
            coef = -c.rho0 * (-pj.w) * dpi;
            if (!TEMP && c.has_dynamic_rigid) {
                const int j = to_global(win, view, idx);
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
}

// compute_surface_tension_acceleration (base_solver.py:209-229): a_i += sum.  aux.w = m_j
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_surface_tension(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(2, d.pv, d.aux);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float sm = c.sigma / d.aux[i].w;
    float3 a = make_float3(0.f, 0.f, 0.f);
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
        if (!(pj.w > 0.0f)) return;
        const float w = r2 > c.diameter2 ? kernel_W_q(c, sqrtf(r2) * c.inv_h) : c.w_diameter;
        const float s = sm * view.get(1, idx).w * w;
        a.x = fmaf(-s, R.x, a.x); a.y = fmaf(-s, R.y, a.y); a.z = fmaf(-s, R.z, a.z);
    });
    float4 acc = d.acc[i];
    d.acc[i] = make_float4(acc.x + a.x, acc.y + a.y, acc.z + a.z, 0.f);
}

// compute_viscosity_acceleration_standard (base_solver.py:231-278): a_i += sum / rho0.
// payload: pv, vm (v_j, m_j), aux.z = rho_j
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_viscosity(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(3, d.pv, d.vm, d.aux);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 vi = d.vm[i];
    const float den_i = d.rho[i];
    float3 a = make_float3(0.f, 0.f, 0.f);
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
        const float4 vj = view.get(1, idx);
        const float v_xy = dot3(make_float3(vi.x - vj.x, vi.y - vj.y, vi.z - vj.z), R);
        const float gs = kernel_gradient_scale(c, r2);
        float coef;
        if (pj.w > 0.0f) {
            coef = c.visc_cf * ((vi.w + vj.w) * 0.5f) / view.get(2, idx).z;
        } else {
            coef = c.visc_cb * (c.rho0 * (-pj.w)) / den_i;
        }
        const float s = coef / (r2 + c.visc_eps) * v_xy * gs;
        a.x = fmaf(s, R.x, a.x); a.y = fmaf(s, R.y, a.y); a.z = fmaf(s, R.z, a.z);
        if (!(pj.w > 0.0f) && c.has_dynamic_rigid) {
            const int j = to_global(win, view, idx);
            if (__ldg(d.is_dynamic + j)) {
                float3 force = R * (-s * vi.w * c.inv_rho0);     // -acc * m_i / rho0
                add_wrench(d, __ldg(d.object_id + j), force, f3(pj));
            }
        }
    });
    float4 acc = d.acc[i];
    d.acc[i] = make_float4(acc.x + a.x * c.inv_rho0, acc.y + a.y * c.inv_rho0, acc.z + a.z * c.inv_rho0, 0.f);
}

// DFSPH compute_alpha (DFSPH.py:22-62)
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_alpha(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(1, d.pv);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    float3 grad_i = make_float3(0.f, 0.f, 0.f);
    float sum_k = 0.0f;
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto&, int, float4 pj, float3 R, float r2) {
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
// (compute_kappa_v :132-137 / compute_kappa :217-223, written to the field and to aux) and the
// error sum (compute_density_derivative_error :205-211 / compute_density_error :285-294).
template <bool STAR, bool LIST, bool FUSED>
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_density_change(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(2, d.pv, d.vm);
    float err = 0.0f;
    if (i < c.N) {
        const float4 pi = d.pv[i];
        if (pi.w > 0.0f) {
            const float4 vi = d.vm[i];
            float delta = 0.0f;
            int nn = 0;
            window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
                const float4 vj = view.get(1, idx);
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
            if (FUSED) d.aux[i] = make_float4(kap, kap / rho, rho, vi.w);
        }
    }
    if (FUSED) block_reduce_add(d.red + RED_ERR, (double)err);
}

// DFSPH correct_divergence_step (DFSPH.py:161-202) / correct_density_error_step (:245-283).
// aux = (kappa_j, kappa_j / rho_j, rho_j, m_j)
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_correct(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(2, d.pv, d.aux);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 ai = d.aux[i];
    const float k_i = ai.x, ki_rho = ai.y;
    const float thresh = 1e-5f * c.dt;   // m_eps * dt
    const bool rigid_on = fabsf(k_i) > thresh;
    float3 dv = make_float3(0.f, 0.f, 0.f);
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
        float s;
        if (pj.w > 0.0f) {
            const float4 aj = view.get(1, idx);
            if (!(fabsf(k_i + aj.x) > thresh)) return;
            s = pj.w * kernel_gradient_scale(c, r2) * (ki_rho + aj.y) * c.rho0;
        } else {
            if (!rigid_on) return;
            s = (-pj.w) * kernel_gradient_scale(c, r2) * ki_rho * c.rho0;
            if (c.has_dynamic_rigid) {
                const int j = to_global(win, view, idx);
                if (__ldg(d.is_dynamic + j)) {
                    float3 force = R * (s * c.inv_dt * (pi.w * c.rho0));
                    add_wrench(d, __ldg(d.object_id + j), force, f3(pj));
                }
            }
        }
        dv.x = fmaf(-s, R.x, dv.x); dv.y = fmaf(-s, R.y, dv.y); dv.z = fmaf(-s, R.z, dv.z);
    });
    float4 v = d.vm[i];
    d.vm[i] = make_float4(v.x + dv.x, v.y + dv.y, v.z + dv.z, v.w);
}

// PCISPH compute_density_star (PCISPH.py:32-62): predicted positions, no self term, neighbour
// set from the current positions.  Accumulates sum max(0, rho*/rho0 - 1) into red[RED_ERR].
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_pcisph_density_star(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(2, d.pv, d.x_pred);
    float err = 0.0f;
    if (i < c.N) {
        const float4 pi = d.pv[i];
        if (pi.w > 0.0f) {
            const float4 xi = d.x_pred[i];
            float ret = 0.0f;
            window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3, float) {
                float4 xj = pj.w > 0.0f ? view.get(1, idx) : pj;
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
// returns c' such that A_ij = c' * (R (x) R)   (grad W = gs * R).  aux_j = (., ., rho_j, m_j)
__device__ __forceinline__ float visc_A_scale(const Consts& c, float mi, float den_i, float4 pj, float4 aux_j, float r2) {
    const float gs = kernel_gradient_scale(c, r2);
    float coef;
    if (pj.w > 0.0f) coef = -c.visc_cf * ((mi + aux_j.w) * 0.5f) / aux_j.z;
    else coef = -c.visc_cb * (c.rho0 * (-pj.w)) / den_i;
    return coef / (r2 + c.visc_eps) * gs;
}

// prepare_conjugate_gradient_solver1, the per-particle part (base_solver.py:300-315):
// D_i^-1, b_i and p_i <- x_i
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_prepare1(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(3, d.pv, d.vm, d.aux);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 vi = d.vm[i];
    const float den_i = d.rho[i];
    // ret = -sum A_ij (symmetric in R (x) R): 6 unique entries
    float sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    float3 b = make_float3(0.f, 0.f, 0.f);
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
        const float a = -visc_A_scale(c, vi.w, den_i, pj, view.get(2, idx), r2);   // ret -= A_ij
        sxx = fmaf(a * R.x, R.x, sxx); sxy = fmaf(a * R.x, R.y, sxy); sxz = fmaf(a * R.x, R.z, sxz);
        syy = fmaf(a * R.y, R.y, syy); syz = fmaf(a * R.y, R.z, syz); szz = fmaf(a * R.z, R.z, szz);
        if (!(pj.w > 0.0f)) {   // compute_b_i_task :333-346, rigid neighbours only
            const float4 vj = view.get(1, idx);
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

// compute_Ap (base_solver.py:373-391): Ap_i = p_i + dt/rho0 * D_i^-1 sum_{fluid j} (-A_ij) p_j
template <bool LIST>
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_Ap(Consts c, Dev d, int wmax) {
    WINDOW_PROLOGUE(3, d.pv, d.cg_p, d.aux);
    if (i >= c.N) return;
    const float4 pi = d.pv[i];
    if (!(pi.w > 0.0f)) return;
    const float4 ai = d.aux[i];
    const float mi = ai.w, den_i = ai.z;
    float3 s = make_float3(0.f, 0.f, 0.f);
    window_neighbors<LIST>(c, d, win, i, pi, [&](const auto& view, int idx, float4 pj, float3 R, float r2) {
        if (!(pj.w > 0.0f)) return;
        const float a = -visc_A_scale(c, mi, den_i, pj, view.get(2, idx), r2);
        const float4 pj_cg = view.get(1, idx);
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

#define LAUNCH_PLAIN(...)                                                     \
    do {                                                                      \
        if (h->c.N > 0) {                                                     \
            SphProf _prof(h, #__VA_ARGS__);                                   \
            __VA_ARGS__<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, h->stream>>>(h->c, h->d); \
            h->launches++;                                                    \
        }                                                                     \
    } while (0)

// window kernel: one CTA per chunk, dynamic shared memory = wmax slots x NPAY float4
#define LAUNCH_WIN(NPAY, ...)                                                 \
    do {                                                                      \
        if (h->c.N > 0) {                                                     \
            SphProf _prof(h, #__VA_ARGS__);                                   \
            const size_t smem_ = (size_t)h->wmax * 16 * (NPAY);               \
            set_smem_limit(__VA_ARGS__, smem_);                               \
            __VA_ARGS__<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, smem_, h->stream>>>(h->c, h->d, h->wmax); \
            h->launches++;                                                    \
        }                                                                     \
    } while (0)

#define LAUNCH_LIST(NPAY, kernel, ...)                                        \
    do {                                                                      \
        if (sph_lists_ready(h)) LAUNCH_WIN(NPAY, kernel<__VA_ARGS__ true>);   \
        else LAUNCH_WIN(NPAY, kernel<__VA_ARGS__ false>);                     \
    } while (0)

// (re)build the lists without touching densities when a list consumer finds them stale
bool sph_lists_ready(SphHandle* h) {
    if (!h->lists_enabled) return false;
    if (!h->list_valid) {
        LAUNCH_WIN(1, k_density<false, true>);
        h->list_valid = true;
    }
    return true;
}

static void prep_aux(SphHandle* h, int mode) {
    switch (mode) {
        case AUX_RHO_M: LAUNCH_PLAIN(k_prep_aux<AUX_RHO_M>); break;
        case AUX_KAPPA: LAUNCH_PLAIN(k_prep_aux<AUX_KAPPA>); break;
        case AUX_KAPPA_V: LAUNCH_PLAIN(k_prep_aux<AUX_KAPPA_V>); break;
        case AUX_PRESSURE: LAUNCH_PLAIN(k_prep_aux<AUX_PRESSURE>); break;
    }
}

void sph_launch_rigid_volume(SphHandle* h) { LAUNCH_PLAIN(k_rigid_volume); }
void sph_launch_density(SphHandle* h) {
    if (h->lists_enabled) {
        LAUNCH_WIN(1, k_density<true, true>);
        h->list_valid = true;
    } else {
        LAUNCH_WIN(1, k_density<true, false>);
    }
}
void sph_launch_pressure_accel(SphHandle* h) { prep_aux(h, AUX_PRESSURE); LAUNCH_LIST(2, k_pressure_accel, false, ); }
void sph_launch_temp_pressure_accel(SphHandle* h) { prep_aux(h, AUX_PRESSURE); LAUNCH_LIST(2, k_pressure_accel, true, ); }
void sph_launch_surface_tension(SphHandle* h) { prep_aux(h, AUX_RHO_M); LAUNCH_LIST(2, k_surface_tension, ); }
void sph_launch_viscosity(SphHandle* h, bool aux_ready) {
    if (!aux_ready) prep_aux(h, AUX_RHO_M);
    LAUNCH_LIST(3, k_viscosity, );
}
void sph_launch_dfsph_alpha(SphHandle* h) { LAUNCH_LIST(1, k_dfsph_alpha, ); }
void sph_launch_dfsph_density_derivative(SphHandle* h, bool fused) {
    if (sph_lists_ready(h)) { if (fused) LAUNCH_WIN(2, k_dfsph_density_change<false, true, true>); else LAUNCH_WIN(2, k_dfsph_density_change<false, true, false>); }
    else { if (fused) LAUNCH_WIN(2, k_dfsph_density_change<false, false, true>); else LAUNCH_WIN(2, k_dfsph_density_change<false, false, false>); }
}
void sph_launch_dfsph_density_star(SphHandle* h, bool fused) {
    if (sph_lists_ready(h)) { if (fused) LAUNCH_WIN(2, k_dfsph_density_change<true, true, true>); else LAUNCH_WIN(2, k_dfsph_density_change<true, true, false>); }
    else { if (fused) LAUNCH_WIN(2, k_dfsph_density_change<true, false, true>); else LAUNCH_WIN(2, k_dfsph_density_change<true, false, false>); }
}
// aux_ready: the fused density-change kernel has just written aux = (kappa, kappa/rho, rho, m)
void sph_launch_dfsph_correct_divergence(SphHandle* h, bool aux_ready) {
    if (!aux_ready) prep_aux(h, AUX_KAPPA_V);
    LAUNCH_LIST(2, k_dfsph_correct, );
}
void sph_launch_dfsph_correct_density(SphHandle* h, bool aux_ready) {
    if (!aux_ready) prep_aux(h, AUX_KAPPA);
    LAUNCH_LIST(2, k_dfsph_correct, );
}
void sph_launch_pcisph_density_star(SphHandle* h) { LAUNCH_LIST(2, k_pcisph_density_star, ); }
void sph_launch_cg_prepare1(SphHandle* h) { prep_aux(h, AUX_RHO_M); LAUNCH_LIST(3, k_cg_prepare1, ); }
void sph_launch_cg_Ap(SphHandle* h, bool aux_ready) {   // aux = (., ., rho, m): unchanged inside the CG loop
    if (!aux_ready) prep_aux(h, AUX_RHO_M);
    LAUNCH_LIST(3, k_cg_Ap, );
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
