// sph_stream.cu — streaming (HBM-bound) kernels: integration, boundary clamp, equation of state,
// DFSPH/PCISPH/CG element-wise updates, reductions, and the field <-> dense-buffer converters
// behind sph_get_field / sph_set_field.  One thread per particle, coalesced float4 / scalar SoA.
#include "sph_kernels.h"

namespace {

#define TID_OR_RETURN(n)                                   \
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   \
    if (i >= (n)) return;
// owned rows only (all particles unless the handle is a Z-slab)
#define ROW_TID_OR_RETURN()                                \
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   \
    SPH_ROW_OR_RETURN(c, i);

// compute_gravity_acceleration (base_solver.py:202-207): assignment, fluid only
__global__ void __launch_bounds__(SPH_BLOCK) k_gravity(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    if (d.pv[i].w > 0.0f) d.acc[i] = make_float4(c.gx, c.gy, c.gz, 0.f);
}

// update_fluid_velocity (base_solver.py:642-649)
__global__ void __launch_bounds__(SPH_BLOCK) k_update_velocity(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    if (!(d.pv[i].w > 0.0f)) return;
    float4 v = d.vm[i];
    const float4 a = d.acc[i];
    v.x += c.dt * a.x; v.y += c.dt * a.y; v.z += c.dt * a.z;
    d.vm[i] = v;
}

// update_fluid_position (base_solver.py:651-666) incl. the emitter branch
__global__ void __launch_bounds__(SPH_BLOCK) k_update_position(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    float4 p = d.pv[i];
    const float4 v = d.vm[i];
    if (p.w > 0.0f) {
        p.x += c.dt * v.x; p.y += c.dt * v.y; p.z += c.dt * v.z;
        d.pv[i] = p;
    } else if (p.y > c.g_upper) {
        const int obj = d.object_id[i];
        if (obj < 0 || obj >= SPH_MAX_OBJECTS) return;   // box particles carry id -1
        if (d.object_material[obj] == SPH_MATERIAL_FLUID) {
            p.x += c.dt * v.x; p.y += c.dt * v.y; p.z += c.dt * v.z;
            if (p.y <= c.g_upper) {
                d.material[i] = SPH_MATERIAL_FLUID;
                p.w = fabsf(p.w);
            }
            d.pv[i] = p;
        }
    }
}

// prepare_emitter (base_solver.py:669-677)
__global__ void __launch_bounds__(SPH_BLOCK) k_prepare_emitter(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    float4 p = d.pv[i];
    if (p.w > 0.0f && p.y > c.g_upper) {
        d.material[i] = SPH_MATERIAL_RIGID;
        p.w = -p.w;
        d.pv[i] = p;
    }
}

// enforce_domain_boundary_3D + simulate_collisions (base_solver.py:544-605)
__global__ void __launch_bounds__(SPH_BLOCK) k_boundary(Consts c, Dev d, int particle_type) {
    ROW_TID_OR_RETURN();
    if (!(d.material[i] == particle_type && d.is_dynamic[i])) return;
    float4 p = d.pv[i];
    const float3 pos = f3(p);
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (pos.x > c.dom_x - c.padding) { n.x += 1.0f; p.x = c.dom_x - c.padding; }
    if (pos.x <= c.padding) { n.x += -1.0f; p.x = c.padding; }
    if (pos.y > c.dom_y - c.padding) { n.y += 1.0f; p.y = c.dom_y - c.padding; }
    if (pos.y <= c.padding) { n.y += -1.0f; p.y = c.padding; }
    if (pos.z > c.dom_z - c.padding) { n.z += 1.0f; p.z = c.dom_z - c.padding; }
    if (pos.z <= c.padding) { n.z += -1.0f; p.z = c.padding; }
    const float len = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    if (len > 1e-6f) {
        d.pv[i] = p;
        const float3 e = make_float3(n.x / len, n.y / len, n.z / len);
        float4 v = d.vm[i];
        const float s = 1.5f * (v.x * e.x + v.y * e.y + v.z * e.z);   // (1 + c_f), c_f = 0.5
        v.x -= s * e.x; v.y -= s * e.y; v.z -= s * e.z;
        d.vm[i] = v;
    }
}

// _renew_rigid_particle_state (base_solver.py:615-629)
__global__ void __launch_bounds__(SPH_BLOCK) k_renew_rigid(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    if (!(d.material[i] == SPH_MATERIAL_RIGID && d.is_dynamic[i])) return;
    const int obj = d.object_id[i];
    if (obj < 0 || obj >= SPH_MAX_OBJECTS || !d.rigid_is_dynamic[obj]) return;
    const float* s = d.rigid_state + obj * 24;  // com0(0..2) com(3..5) rot(6..14) vel(15..17) omega(18..20)
    const float3 q = make_float3(d.x0[3 * i] - s[0], d.x0[3 * i + 1] - s[1], d.x0[3 * i + 2] - s[2]);
    const float3 r = make_float3(s[6] * q.x + s[7] * q.y + s[8] * q.z, s[9] * q.x + s[10] * q.y + s[11] * q.z,
                                 s[12] * q.x + s[13] * q.y + s[14] * q.z);
    float4 p = d.pv[i];
    p.x = s[3] + r.x; p.y = s[4] + r.y; p.z = s[5] + r.z;
    d.pv[i] = p;
    const float3 w = make_float3(s[18], s[19], s[20]);
    const float3 wxr = cross3(w, r);
    float4 v = d.vm[i];
    v.x = s[15] + wxr.x; v.y = s[16] + wxr.y; v.z = s[17] + wxr.z;
    d.vm[i] = v;
}

// WCSPHSolver.compute_pressure (WCSPH.py:16-24): clamp written back, Tait EOS gamma = 7, B = 5e4
__global__ void __launch_bounds__(SPH_BLOCK) k_wcsph_pressure(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    if (!(d.pv[i].w > 0.0f)) return;
    const float rho = fmaxf(d.rho[i], c.rho0);
    d.rho[i] = rho;
    d.p[i] = 50000.0f * (powf(rho / c.rho0, 7.0f) - 1.0f);
}

// DFSPH compute_kappa_v (DFSPH.py:132-137) / compute_kappa (:217-223)
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_kappa_v(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    if (d.pv[i].w > 0.0f) d.kappa_v[i] = d.drho[i] * d.alpha[i];
}
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_kappa(Consts c, Dev d) {
    ROW_TID_OR_RETURN();
    if (d.pv[i].w > 0.0f) d.kappa[i] = (d.rho_star[i] - 1.0f) * d.alpha[i] * c.inv_dt;
}

// compute_density_derivative_error (DFSPH.py:205-211) / compute_density_error (:285-294): sums only;
// the host divides by particle_num (all particles, App. B#7)
template <bool DIVERGENCE>
__global__ void __launch_bounds__(SPH_BLOCK) k_dfsph_error(Consts c, Dev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float e = 0.0f;
    if (SPH_IS_ROW(c, i) && d.pv[i].w > 0.0f) e = DIVERGENCE ? c.rho0 * d.drho[i] : d.rho_star[i] - 1.0f;
    block_reduce_add(d.red + RED_ERR, (double)e);
}

// PCISPH streaming kernels (PCISPH.py:18-29, 65-71, 153-162)
__global__ void __launch_bounds__(SPH_BLOCK) k_pcisph_predict_velocity(Consts c, Dev d) {
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    const float4 v = d.vm[i], a = d.acc[i], ap = d.a_p[i];
    d.v_pred[i] = make_float4(v.x + c.dt * (a.x + ap.x), v.y + c.dt * (a.y + ap.y), v.z + c.dt * (a.z + ap.z), 0.f);
}
__global__ void __launch_bounds__(SPH_BLOCK) k_pcisph_predict_position(Consts c, Dev d) {
    TID_OR_RETURN(c.N);
    const float4 p = d.pv[i];
    if (!(p.w > 0.0f)) return;
    const float4 v = d.v_pred[i];
    d.x_pred[i] = make_float4(p.x + c.dt * v.x, p.y + c.dt * v.y, p.z + c.dt * v.z, 0.f);
}
__global__ void __launch_bounds__(SPH_BLOCK) k_pcisph_update_pressure(Consts c, Dev d) {
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    float p = d.p[i] + c.pcisph_k * (c.rho0 - d.rho_star[i]);
    d.p[i] = p < 0.0f ? 0.0f : p;
}
__global__ void __launch_bounds__(SPH_BLOCK) k_pcisph_init_step(Consts c, Dev d) {
    TID_OR_RETURN(c.cap);
    d.a_p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    d.p[i] = 0.0f;
    if (i >= c.N) return;
    const float4 p = d.pv[i];
    if (!(p.w > 0.0f)) return;
    const float4 v = d.vm[i], a = d.acc[i];
    const float4 vp = make_float4(v.x + c.dt * a.x, v.y + c.dt * a.y, v.z + c.dt * a.z, 0.f);
    d.v_pred[i] = vp;
    d.x_pred[i] = make_float4(p.x + c.dt * vp.x, p.y + c.dt * vp.y, p.z + c.dt * vp.z, 0.f);
}

// ---- implicit viscosity CG element-wise parts (base_solver.py:281-473) ----
__device__ __forceinline__ float cg_alpha_of(const Dev& d) {
    const float num = (float)d.red[RED_CG_RR], den = (float)d.red[RED_CG_PAP];
    return den > 1e-18f ? num / den : 0.0f;
}
__device__ __forceinline__ float cg_beta_of(const Dev& d) {
    const float num = (float)d.red[RED_CG_RR_NEW], den = (float)d.red[RED_CG_RR_OLD];
    return den > 1e-18f ? num / den : 0.0f;
}
// prepare_conjugate_gradient_solver1, the fills + warm start (:284-298)
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_prepare1_pre(Consts c, Dev d) {
    TID_OR_RETURN(c.cap);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    d.cg_r[i] = z; d.cg_p[i] = z; d.cg_b[i] = z; d.cg_Ap[i] = z;
    float4 vo = z;
    if (i < c.N && d.pv[i].w > 0.0f) {
        const float4 v = d.vm[i];
        float4 x = d.cg_x[i];
        d.cg_x[i] = make_float4(x.x + v.x, x.y + v.y, x.z + v.z, 0.f);
        vo = make_float4(v.x, v.y, v.z, 0.f);
    }
    d.v_orig[i] = vo;
}
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_prepare2(Consts c, Dev d) {   // :317-323
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    const float* m = d.cg_dinv + 9 * (size_t)i;
    const float4 b = d.cg_b[i], ap = d.cg_Ap[i];
    const float4 r = make_float4(m[0] * b.x + m[1] * b.y + m[2] * b.z - ap.x, m[3] * b.x + m[4] * b.y + m[5] * b.z - ap.y,
                                 m[6] * b.x + m[7] * b.y + m[8] * b.z - ap.z, 0.f);
    d.cg_r[i] = r;
    d.cg_p[i] = r;
}
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_dots(Consts c, Dev d) {   // compute_cg_alpha :393-406
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float rr = 0.f, pap = 0.f;
    if (i < c.N && d.pv[i].w > 0.0f) {
        const float4 r = d.cg_r[i], p = d.cg_p[i], ap = d.cg_Ap[i];
        rr = r.x * r.x + r.y * r.y + r.z * r.z;
        pap = p.x * ap.x + p.y * ap.y + p.z * ap.z;
    }
    block_reduce_add(d.red + RED_CG_RR, (double)rr);
    block_reduce_add(d.red + RED_CG_PAP, (double)pap);
}
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_update_x(Consts c, Dev d) {   // :408-412
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    const float alpha = cg_alpha_of(d);
    const float4 p = d.cg_p[i];
    float4 x = d.cg_x[i];
    d.cg_x[i] = make_float4(x.x + alpha * p.x, x.y + alpha * p.y, x.z + alpha * p.z, 0.f);
}
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_update_r(Consts c, Dev d) {   // update_cg_r_and_beta :414-431
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float nn = 0.f, oo = 0.f;
    if (i < c.N && d.pv[i].w > 0.0f) {
        const float alpha = cg_alpha_of(d);
        const float4 r = d.cg_r[i], ap = d.cg_Ap[i];
        const float4 nr = make_float4(r.x - alpha * ap.x, r.y - alpha * ap.y, r.z - alpha * ap.z, 0.f);
        nn = nr.x * nr.x + nr.y * nr.y + nr.z * nr.z;
        oo = r.x * r.x + r.y * r.y + r.z * r.z;
        d.cg_r[i] = nr;
    }
    block_reduce_add(d.red + RED_CG_RR_NEW, (double)nn);
    block_reduce_add(d.red + RED_CG_RR_OLD, (double)oo);
}
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_update_p(Consts c, Dev d) {   // :433-437
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    const float beta = cg_beta_of(d);
    const float4 r = d.cg_r[i], p = d.cg_p[i];
    d.cg_p[i] = make_float4(r.x + beta * p.x, r.y + beta * p.y, r.z + beta * p.z, 0.f);
}
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_prepare_guess(Consts c, Dev d) {   // :439-443
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    const float4 x = d.cg_x[i], vo = d.v_orig[i];
    d.cg_x[i] = make_float4(x.x - vo.x, x.y - vo.y, x.z - vo.z, 0.f);
}
template <bool RESTORE>
__global__ void __launch_bounds__(SPH_BLOCK) k_cg_velocity(Consts c, Dev d) {   // :463-473
    TID_OR_RETURN(c.N);
    if (!(d.pv[i].w > 0.0f)) return;
    const float4 s = RESTORE ? d.v_orig[i] : d.cg_x[i];
    float4 v = d.vm[i];
    d.vm[i] = make_float4(s.x, s.y, s.z, v.w);
}

// compute_rigid_body_mass (base_container.py:384-390)
__global__ void __launch_bounds__(SPH_BLOCK) k_rigid_body_mass(Consts c, Dev d, int object_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (i < c.N && d.object_id[i] == object_id && d.is_dynamic[i]) m = d.rho[i] * c.V0;
    block_reduce_add(d.red + RED_MASS, (double)m);
}

__global__ void __launch_bounds__(SPH_BLOCK) k_count_dynamic_rigid(Consts c, Dev d, int* out) {
    TID_OR_RETURN(c.N);
    if (d.material[i] == SPH_MATERIAL_RIGID && d.is_dynamic[i]) atomicOr(out, 1);
}

__global__ void k_fill_i32(int* p, size_t n, int v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_f32(float* p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- field <-> dense staging -------------------------------------------------------------------
enum Conv { CONV_POS, CONV_VEL, CONV_VOLUME, CONV_MASS, CONV_MATERIAL, CONV_F4 };

template <int CONV, bool TO_STAGING>
__global__ void __launch_bounds__(SPH_BLOCK) k_convert(Dev d, float4* f4, float* staging_f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int* staging_i = reinterpret_cast<int*>(staging_f);
    if (CONV == CONV_POS || CONV == CONV_VEL || CONV == CONV_F4) {
        float4* a = CONV == CONV_POS ? d.pv : (CONV == CONV_VEL ? d.vm : f4);
        if (TO_STAGING) {
            const float4 v = a[i];
            staging_f[3 * i] = v.x; staging_f[3 * i + 1] = v.y; staging_f[3 * i + 2] = v.z;
        } else {
            float4 v = a[i];
            v.x = staging_f[3 * i]; v.y = staging_f[3 * i + 1]; v.z = staging_f[3 * i + 2];
            a[i] = v;
        }
    } else if (CONV == CONV_VOLUME) {
        if (TO_STAGING) staging_f[i] = fabsf(d.pv[i].w);
        else {
            float4 v = d.pv[i];
            const float a = fabsf(staging_f[i]);
            v.w = d.material[i] == SPH_MATERIAL_FLUID ? a : -a;
            d.pv[i] = v;
        }
    } else if (CONV == CONV_MASS) {
        if (TO_STAGING) staging_f[i] = d.vm[i].w;
        else reinterpret_cast<float*>(d.vm + i)[3] = staging_f[i];
    } else if (CONV == CONV_MATERIAL) {
        if (TO_STAGING) staging_i[i] = d.material[i];
        else {
            const int m = staging_i[i];
            d.material[i] = m;
            float4 v = d.pv[i];
            const float a = fabsf(v.w);
            v.w = m == SPH_MATERIAL_FLUID ? a : -a;
            d.pv[i] = v;
        }
    }
}

__global__ void __launch_bounds__(SPH_BLOCK) k_cell_coords(Consts c, Dev d, int* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = d.pv[i];
    // unclamped trunc(x / h), exactly pos_to_index (base_container.py:467-469)
    out[3 * i] = (int)(p.x / c.h); out[3 * i + 1] = (int)(p.y / c.h); out[3 * i + 2] = (int)(p.z / c.h);
}

// histogram in the reference's z-fastest flatten (base_container.py:472-481), for grid_num_particles
__global__ void __launch_bounds__(SPH_BLOCK) k_ref_cell_hist(Consts c, Dev d, int* hist) {
    TID_OR_RETURN(c.N);
    const float4 p = d.pv[i];
    const int3 g = cell_of(c, p.x, p.y, p.z);
    atomicAdd(hist + (g.x * c.ny + g.y) * c.nz + g.z, 1);
}

}  // namespace

#define GRID(n) ((int)(((size_t)(n) + SPH_BLOCK - 1) / SPH_BLOCK))
#define LAUNCH_N(kernel, n, ...)                                                  \
    do {                                                                          \
        if ((n) > 0) {                                                            \
            SphProf _prof(h, #kernel);                                            \
            kernel<<<GRID(n), SPH_BLOCK, 0, h->stream>>>(__VA_ARGS__);            \
            h->launches++;                                                        \
        }                                                                         \
    } while (0)

void sph_launch_gravity(SphHandle* h) { LAUNCH_N(k_gravity, h->c.N, h->c, h->d); }
void sph_launch_update_velocity(SphHandle* h) { LAUNCH_N(k_update_velocity, h->c.N, h->c, h->d); sph_ghost_dirty(h, GHOST_VEL); }
void sph_launch_update_position(SphHandle* h) { LAUNCH_N(k_update_position, h->c.N, h->c, h->d); h->list_valid = false; sph_ghost_dirty(h, GHOST_PV); }
void sph_launch_boundary(SphHandle* h, int t) { LAUNCH_N(k_boundary, h->c.N, h->c, h->d, t); h->list_valid = false; sph_ghost_dirty(h, GHOST_PV | GHOST_VEL); }
void sph_launch_renew_rigid(SphHandle* h) { LAUNCH_N(k_renew_rigid, h->c.N, h->c, h->d); h->list_valid = false; sph_ghost_dirty(h, GHOST_PV | GHOST_VEL); }
void sph_launch_prepare_emitter(SphHandle* h) { LAUNCH_N(k_prepare_emitter, h->c.N, h->c, h->d); h->list_valid = false; sph_ghost_dirty(h, GHOST_PV); }
void sph_launch_wcsph_pressure(SphHandle* h) { LAUNCH_N(k_wcsph_pressure, h->c.N, h->c, h->d); sph_ghost_dirty(h, GHOST_RHO); }
void sph_launch_dfsph_kappa_v(SphHandle* h) { LAUNCH_N(k_dfsph_kappa_v, h->c.N, h->c, h->d); }
void sph_launch_dfsph_kappa(SphHandle* h) { LAUNCH_N(k_dfsph_kappa, h->c.N, h->c, h->d); }
void sph_launch_dfsph_divergence_error(SphHandle* h) { LAUNCH_N(k_dfsph_error<true>, h->c.N, h->c, h->d); }
void sph_launch_dfsph_density_error(SphHandle* h) { LAUNCH_N(k_dfsph_error<false>, h->c.N, h->c, h->d); }
void sph_launch_pcisph_predict_velocity(SphHandle* h) { LAUNCH_N(k_pcisph_predict_velocity, h->c.N, h->c, h->d); }
void sph_launch_pcisph_predict_position(SphHandle* h) { LAUNCH_N(k_pcisph_predict_position, h->c.N, h->c, h->d); }
void sph_launch_pcisph_update_pressure(SphHandle* h) { LAUNCH_N(k_pcisph_update_pressure, h->c.N, h->c, h->d); }
void sph_launch_pcisph_init_step(SphHandle* h) { LAUNCH_N(k_pcisph_init_step, h->c.cap, h->c, h->d); }
void sph_launch_cg_prepare1_pre(SphHandle* h) { LAUNCH_N(k_cg_prepare1_pre, h->c.cap, h->c, h->d); }
void sph_launch_cg_prepare2(SphHandle* h) { LAUNCH_N(k_cg_prepare2, h->c.N, h->c, h->d); }
void sph_launch_cg_dots(SphHandle* h) { LAUNCH_N(k_cg_dots, h->c.N, h->c, h->d); }
void sph_launch_cg_update_x(SphHandle* h) { LAUNCH_N(k_cg_update_x, h->c.N, h->c, h->d); }
void sph_launch_cg_update_r(SphHandle* h) { LAUNCH_N(k_cg_update_r, h->c.N, h->c, h->d); }
void sph_launch_cg_update_p(SphHandle* h) { LAUNCH_N(k_cg_update_p, h->c.N, h->c, h->d); }
void sph_launch_cg_prepare_guess(SphHandle* h) { LAUNCH_N(k_cg_prepare_guess, h->c.N, h->c, h->d); }
void sph_launch_cg_velocity_from_x(SphHandle* h) { LAUNCH_N(k_cg_velocity<false>, h->c.N, h->c, h->d); sph_ghost_dirty(h, GHOST_VEL); }
void sph_launch_cg_velocity_restore(SphHandle* h) { LAUNCH_N(k_cg_velocity<true>, h->c.N, h->c, h->d); sph_ghost_dirty(h, GHOST_VEL); }
void sph_launch_rigid_body_mass(SphHandle* h, int obj) { LAUNCH_N(k_rigid_body_mass, h->c.N, h->c, h->d, obj); }
void sph_launch_count_dynamic_rigid(SphHandle* h, int* out) { LAUNCH_N(k_count_dynamic_rigid, h->c.N, h->c, h->d, out); }
void sph_fill_i32(SphHandle* h, int* p, size_t n, int v) { LAUNCH_N(k_fill_i32, n, p, n, v); }
void sph_fill_f32(SphHandle* h, float* p, size_t n, float v) { LAUNCH_N(k_fill_f32, n, p, n, v); }
void sph_launch_cell_coords(SphHandle* h, int* out, int n) { LAUNCH_N(k_cell_coords, n, h->c, h->d, out, n); }
void sph_launch_ref_cell_hist(SphHandle* h, int* hist) { LAUNCH_N(k_ref_cell_hist, h->c.N, h->c, h->d, hist); }

static float4* f4_field(SphHandle* h, int field) {
    Dev& d = h->d;
    switch (field) {
        case SPH_F_ACCELERATION: return d.acc;
        case SPH_F_PRESSURE_ACCELERATION: return d.a_p;
        case SPH_F_PREDICTED_VELOCITY: return d.v_pred;
        case SPH_F_PREDICTED_POSITION: return d.x_pred;
        case SPH_F_CG_P: return d.cg_p;
        case SPH_F_ORIGINAL_VELOCITY: return d.v_orig;
        case SPH_F_CG_AP: return d.cg_Ap;
        case SPH_F_CG_X: return d.cg_x;
        case SPH_F_CG_B: return d.cg_b;
        case SPH_F_CG_R: return d.cg_r;
        default: return nullptr;
    }
}

// dense 4-byte-word arrays that are stored exactly as the host sees them
static void* dense_field(SphHandle* h, int field, int* comps) {
    Dev& d = h->d;
    *comps = 1;
    switch (field) {
        case SPH_F_OBJECT_ID: return d.object_id;
        case SPH_F_DENSITY: return d.rho;
        case SPH_F_PRESSURE: return d.p;
        case SPH_F_COLOR: *comps = 3; return d.color;
        case SPH_F_IS_DYNAMIC: return d.is_dynamic;
        case SPH_F_ORIGINAL_POSITION: *comps = 3; return d.x0;
        case SPH_F_GRID_ID: return d.grid_id;
        case SPH_F_UID: return d.uid;
        case SPH_F_DFSPH_ALPHA: return d.alpha;
        case SPH_F_DFSPH_KAPPA: return d.kappa;
        case SPH_F_DFSPH_KAPPA_V: return d.kappa_v;
        case SPH_F_DENSITY_STAR: return d.rho_star;
        case SPH_F_DENSITY_DERIVATIVE: return d.drho;
        case SPH_F_CG_DIAG_INV: *comps = 9; return d.cg_dinv;
        default: return nullptr;
    }
}

// returns components (>0) on success; the dense copy of field[0..n) is in h->staging
int sph_field_to_staging(SphHandle* h, int field, int n) {
    float* st = (float*)h->staging;
    int comps = 0;
    if (void* p = dense_field(h, field, &comps)) {
        if (!p) return SPH_E_STATE;
        cudaMemcpyAsync(st, p, (size_t)n * comps * 4, cudaMemcpyDeviceToDevice, h->stream);
        return comps;
    }
    switch (field) {
        case SPH_F_POSITION: LAUNCH_N((k_convert<CONV_POS, true>), n, h->d, nullptr, st, n); return 3;
        case SPH_F_VELOCITY: LAUNCH_N((k_convert<CONV_VEL, true>), n, h->d, nullptr, st, n); return 3;
        case SPH_F_REST_VOLUME: LAUNCH_N((k_convert<CONV_VOLUME, true>), n, h->d, nullptr, st, n); return 1;
        case SPH_F_MASS: LAUNCH_N((k_convert<CONV_MASS, true>), n, h->d, nullptr, st, n); return 1;
        case SPH_F_MATERIAL: LAUNCH_N((k_convert<CONV_MATERIAL, true>), n, h->d, nullptr, st, n); return 1;
        default: break;
    }
    if (float4* f = f4_field(h, field)) {
        LAUNCH_N((k_convert<CONV_F4, true>), n, h->d, f, st, n);
        return 3;
    }
    return SPH_E_INVALID;
}

int sph_staging_to_field(SphHandle* h, int field, int n) {
    float* st = (float*)h->staging;
    int comps = 0;
    if (void* p = dense_field(h, field, &comps)) {
        cudaMemcpyAsync(p, st, (size_t)n * comps * 4, cudaMemcpyDeviceToDevice, h->stream);
        return comps;
    }
    switch (field) {
        case SPH_F_POSITION: LAUNCH_N((k_convert<CONV_POS, false>), n, h->d, nullptr, st, n); return 3;
        case SPH_F_VELOCITY: LAUNCH_N((k_convert<CONV_VEL, false>), n, h->d, nullptr, st, n); return 3;
        case SPH_F_REST_VOLUME: LAUNCH_N((k_convert<CONV_VOLUME, false>), n, h->d, nullptr, st, n); return 1;
        case SPH_F_MASS: LAUNCH_N((k_convert<CONV_MASS, false>), n, h->d, nullptr, st, n); return 1;
        case SPH_F_MATERIAL: LAUNCH_N((k_convert<CONV_MATERIAL, false>), n, h->d, nullptr, st, n); return 1;
        default: break;
    }
    if (float4* f = f4_field(h, field)) {
        LAUNCH_N((k_convert<CONV_F4, false>), n, h->d, f, st, n);
        return 3;
    }
    return SPH_E_INVALID;
}
