// sph_api.cu — the C ABI of include/sph_b200.h over the CUDA kernels: handle lifetime, host <->
// device field access, one entry per upstream kernel (sph_run_task), the solver loops and the
// whole-step driver.  Nothing here throws across the ABI; CUDA errors become SPH_E_CUDA with the
// runtime's message in sph_last_error().
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sph_kernels.h"
#include "sph_brick.cuh"

namespace {

int fail(SphHandle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}

int check_cuda(SphHandle* h, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return SPH_OK;
    if (h) h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return SPH_E_CUDA;
}

#define CUDA_TRY(h, expr)                                 \
    do {                                                  \
        int _rc = check_cuda((h), (expr), #expr);         \
        if (_rc) return _rc;                              \
    } while (0)

// every ABI entry runs on the handle's device whatever the caller's current device is
#define ON_DEVICE(h)                                                                             \
    do {                                                                                         \
        if (cudaSetDevice((h)->P.device) != cudaSuccess) return check_cuda((h), cudaGetLastError(), "cudaSetDevice"); \
    } while (0)

int last_launch(SphHandle* h) {
    if (h->sticky_rc) {   // an NCCL halo inside a launcher failed; message is already in h->err
        const int rc = h->sticky_rc;
        h->sticky_rc = 0;
        return rc;
    }
    return check_cuda(h, cudaGetLastError(), "kernel launch");
}

// rows the kernels update: everything, unless the handle is a Z-slab (then set by the sort)
void rows_all(SphHandle* h) {
    if (!h->slab) {
        h->c.row_begin = 0;
        h->c.row_end = h->c.N;
    } else if (!h->rows_from_sort) {
        h->c.row_begin = 0;          // before the first exchange every local particle is owned
        h->c.row_end = h->c.N;
    }
}
// particle_num of the whole domain (DFSPH error normalisation, DFSPH.py:211,294)
float global_particle_num(const SphHandle* h) { return h->slab ? (float)h->n_global : (float)h->c.N; }

template <class T>
int dev_alloc(SphHandle* h, T*& p, size_t count) {
    void* q = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(T) + 256;   // slack so 16-byte tail reads stay in bounds
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return check_cuda(h, e, "cudaMalloc");
    e = cudaMemsetAsync(q, 0, bytes, h->stream);
    if (e != cudaSuccess) return check_cuda(h, e, "cudaMemset");
    h->allocations.push_back(q);
    p = (T*)q;
    return SPH_OK;
}

// smallest float t with sqrtf(t) >= h, so that (r2 < t) == (sqrtf(r2) < h) for every float r2
float neighbor_threshold(float h) {
    float t = h * h;
    while (sqrtf(t) >= h) t = nextafterf(t, 0.0f);
    while (sqrtf(t) < h) t = nextafterf(t, INFINITY);
    return t;
}

void refresh_consts(SphHandle* h) {
    const SphParams& P = h->P;
    Consts& c = h->c;
    c.h = (float)P.dh;
    c.inv_h = 1.0f / c.h;
    c.h2_thresh = neighbor_threshold(c.h);
    const double k = 8.0 / M_PI / (P.dh * P.dh * P.dh);
    c.kW = (float)k; c.kW2 = (float)(2.0 * k); c.kG = (float)(6.0 * k);
    c.kG_inv_h = (float)(6.0 * k / P.dh);
    c.V0 = (float)P.V0; c.rho0 = (float)P.density0; c.inv_rho0 = 1.0f / c.rho0;
    c.dt = (float)P.dt; c.inv_dt = 1.0f / c.dt;
    c.corr_thresh = 1e-5f * c.dt;
    c.g_upper = (float)P.g_upper;
    c.gx = (float)P.gravity[0]; c.gy = (float)P.gravity[1]; c.gz = (float)P.gravity[2];
    c.visc_cf = (float)(2.0 * (3 + 2) * P.viscosity);
    c.visc_cb = (float)(2.0 * (3 + 2) * P.viscosity_b);
    c.visc_eps = (float)(0.01 * P.dh * P.dh);
    c.sigma = (float)P.surface_tension;
    c.diameter = (float)(2.0 * P.dx);
    c.diameter2 = (float)(2.0 * P.dx * 2.0 * P.dx);
    {   // W(diameter) with the same f32 steps as kernel_W (base_solver.py:229)
        float q = c.diameter / c.h;
        float w = 0.f;
        if (q <= 1.0f) w = q <= 0.5f ? c.kW * (6.0f * q * q * q - 6.0f * q * q + 1.0f) : c.kW2 * powf(1.0f - q, 3.0f);
        c.w_diameter = w;
    }
    c.dom_x = (float)P.domain_size[0]; c.dom_y = (float)P.domain_size[1]; c.dom_z = (float)P.domain_size[2];
    c.padding = (float)P.padding;
}

int read_red(SphHandle* h) {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_red, h->d.red, sizeof(double) * RED_COUNT, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SPH_OK;
}
int zero_red(SphHandle* h, int slot, int count = 1) {
    CUDA_TRY(h, cudaMemsetAsync(h->d.red + slot, 0, sizeof(double) * count, h->stream));
    return SPH_OK;
}

int update_dynamic_rigid_flag(SphHandle* h) {
    if (!h->dyn_rigid_dirty) return SPH_OK;
    if (sph_is_slab(h)) return SPH_OK;   // Z-slabs: the flag is global, summed over the ranks by the next sort (sph_slab.cu)
    int* flag = h->d.scan_tmp;   // any int scratch
    CUDA_TRY(h, cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
    sph_launch_count_dynamic_rigid(h, flag);
    int v = 0;
    CUDA_TRY(h, cudaMemcpyAsync(&v, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->c.has_dynamic_rigid = v;
    h->dyn_rigid_dirty = false;
    return SPH_OK;
}

// compute_pcisph_k (PCISPH.py:128-151), a 7^3 template lattice: evaluated on the host in f32
float pcisph_k_host(const SphHandle* h) {
    const Consts& c = h->c;
    const float diam = (float)(h->P.dx * 2.0 * 0.97);
    const int max_i = (int)(c.h / diam) + 1;
    float sx = 0, sy = 0, sz = 0, s2 = 0;
    for (int i = -max_i; i <= max_i; i++)
        for (int j = -max_i; j <= max_i; j++)
            for (int k = -max_i; k <= max_i; k++) {
                const float x = -(i * diam), y = -(j * diam), z = -(k * diam);
                const float r2 = fmaf(z, z, fmaf(y, y, x * x));
                const float r = sqrtf(r2);
                if (!(r < c.h)) continue;
                const float q = r / c.h;
                float gs = 0.f;
                if (r > 1e-5f && q <= 1.0f) gs = (q <= 0.5f ? c.kG * q * (3.0f * q - 2.0f) : -c.kG * (1.0f - q) * (1.0f - q)) / (r * c.h);
                const float gx = gs * x, gy = gs * y, gz = gs * z;
                sx += gx; sy += gy; sz += gz;
                s2 += gx * gx + gy * gy + gz * gz;
            }
    return -0.5f / (c.dt * c.V0) / (c.dt * c.V0) / (sx * sx + sy * sy + sz * sz + s2);
}

// ---- solver loops ---------------------------------------------------------------------------
// DFSPH.correct_divergence_error (DFSPH.py:139-159) / correct_density_error (:225-243).
// Density change, the kappa of the next correction step and the error sum are one fused kernel (same arithmetic as the
// three upstream kernels; sph_run_task exposes them separately).  The loop-exit test runs on the device, in the last
// CTA of every density-change launch; the host launches a batch of iterations back to back — as many as the previous
// solve of this kind needed — and reads the outcome once per batch.  Launches after convergence return at once, so the
// solve stops at the same iteration as the reference's host-driven loop.
int dfsph_solve(SphHandle* h, bool density, int* iters, float* err) {
    const Consts& c = h->c;
    const float eta = density ? 0.0001f : 0.001f * c.rho0 / c.dt;
    int rc;
    auto density_change = [&](int mode, bool spec) {
        if (density) sph_launch_dfsph_density_star(h, true, mode, spec, eta);
        else sph_launch_dfsph_density_derivative(h, true, mode, spec, eta);
    };
    auto correct = [&](bool spec) {
        if (density) sph_launch_dfsph_correct_density(h, true, spec);
        else sph_launch_dfsph_correct_divergence(h, true, spec);
    };
    int it = 0;
    float e = 0.f;
    // Z-slabs: the error sum crosses ranks, so an iteration is correction, density change (error sum only), all-reduce
    // and the exit test as a one-thread kernel; every rank reads the same CTRL words and launches the same batches
    const bool slab = sph_is_slab(h);
    if ((rc = zero_red(h, RED_ERR))) return rc;
    if ((rc = zero_red(h, CTRL_DONE, CTRL_COUNT))) return rc;
    // with peer memory mapped, the ghost refreshes and the error sum of the loop go through the neighbours' memory
    h->peer_loop = slab && sph_slab_peers_ready(h);
    density_change(0, false);
    int& hint = h->solve_hint[density ? 0 : 1];
    int launched = 0;
    double ctrl[CTRL_COUNT] = {0.0, 0.0, 0.0};
    while (launched < 1000) {
        int batch = launched == 0 ? (hint < 1 ? 1 : hint) : (hint / 8 < 2 ? 2 : hint / 8);
        if (h->solve_batch > 0) batch = h->solve_batch;   // SPH_B200_BATCH_ITERS: fixed batch size (1 = a host read per iteration)
        if (batch > 1000 - launched) batch = 1000 - launched;
        for (int b = 0; b < batch; b++) {
            correct(true);
            if (!slab) {
                density_change(2, true);
            } else {
                density_change(1, true);   // peer loop: the epilogue delivers the error sum to every rank
                if (!h->peer_loop && (rc = sph_slab_allreduce_red(h, RED_ERR, 1))) { h->peer_loop = false; return rc; }
                sph_launch_dfsph_solve_check(h, eta);
            }
        }
        launched += batch;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_red + CTRL_DONE, h->d.red + CTRL_DONE, sizeof(double) * CTRL_COUNT, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (int k = 0; k < CTRL_COUNT; k++) ctrl[k] = h->h_red[CTRL_DONE + k];
        if (ctrl[0] != 0.0) break;
    }
    h->peer_loop = false;
    h->peer_signalled = 0;   // what is still stale after the loop is refreshed through NCCL
    it = (int)ctrl[CTRL_ITERS - CTRL_DONE];
    e = (float)ctrl[CTRL_ERR - CTRL_DONE];
    hint = it;
    *iters = it; *err = e;
    return last_launch(h);
}
int dfsph_correct_divergence_error(SphHandle* h, int* iters, float* err) { return dfsph_solve(h, false, iters, err); }
int dfsph_correct_density_error(SphHandle* h, int* iters, float* err) { return dfsph_solve(h, true, iters, err); }

int pcisph_density_star(SphHandle* h) {
    int rc;
    if ((rc = zero_red(h, RED_ERR))) return rc;
    sph_launch_pcisph_density_star(h);
    return SPH_OK;
}
int pcisph_fetch_error(SphHandle* h) {
    int rc;
    if ((rc = read_red(h))) return rc;
    h->density_error = h->Nfluid > 0 ? (float)h->h_red[RED_ERR] / (float)h->Nfluid : 0.0f;
    return SPH_OK;
}

// PCISPH.refine (PCISPH.py:110-125)
int pcisph_refine(SphHandle* h, int* iters, float* err) {
    int it = 0, rc;
    while (it < 1000) {
        if ((rc = pcisph_density_star(h))) return rc;
        sph_launch_pcisph_update_pressure(h);
        sph_launch_temp_pressure_accel(h);
        sph_launch_pcisph_predict_velocity(h);
        sph_launch_pcisph_predict_position(h);
        if ((rc = pcisph_fetch_error(h))) return rc;
        it++;
        if (h->density_error < 0.001f) break;
    }
    *iters = it; *err = h->density_error;
    return last_launch(h);
}

int cg_update_r(SphHandle* h) {
    int rc;
    if ((rc = zero_red(h, RED_CG_RR_NEW, 2))) return rc;
    sph_launch_cg_update_r(h);
    if ((rc = read_red(h))) return rc;
    h->cg_error = sqrtf((float)h->h_red[RED_CG_RR_NEW]);
    return SPH_OK;
}
int cg_dots(SphHandle* h) {
    int rc;
    if ((rc = zero_red(h, RED_CG_RR, 2))) return rc;
    sph_launch_cg_dots(h);
    return SPH_OK;
}

// BaseSolver.implicit_viscosity_solve (base_solver.py:509-517) with conjugate_gradient_loop (:445-461)
int implicit_viscosity_solve(SphHandle* h, int* iters, float* err) {
    int rc;
    sph_launch_cg_prepare1_pre(h);
    sph_launch_cg_prepare1(h);
    sph_launch_cg_Ap(h, true);
    sph_launch_cg_prepare2(h);
    float tol = 1000.0f;
    int it = 0;
    while (tol > 1e-6f && it < 1000) {
        sph_launch_cg_Ap(h, true);
        if ((rc = cg_dots(h))) return rc;
        sph_launch_cg_update_x(h);
        if ((rc = cg_update_r(h))) return rc;
        sph_launch_cg_update_p(h);
        tol = h->cg_error;
        it++;
    }
    sph_launch_cg_velocity_from_x(h);
    sph_launch_viscosity(h);
    sph_launch_cg_velocity_restore(h);
    sph_launch_cg_prepare_guess(h);
    *iters = it; *err = tol;
    return last_launch(h);
}

// compute_non_pressure_acceleration (base_solver.py:190-200)
int non_pressure_acceleration(SphHandle* h, SphStepStats* st) {
    sph_launch_gravity(h);
    sph_launch_surface_tension(h);
    if (h->P.visc_method == SPH_VISC_STANDARD) {
        sph_launch_viscosity(h);
    } else {
        int it; float e;
        int rc = implicit_viscosity_solve(h, &it, &e);
        if (rc) return rc;
        st->cg_iterations = it; st->cg_error = e; st->total_cg_iterations += it;
    }
    return SPH_OK;
}

int step_once(SphHandle* h, SphStepStats* st) {
    int rc, it; float e;
    switch (h->P.method) {
        case SPH_METHOD_WCSPH:   // WCSPH.py:27-45
            if ((rc = sph_sort_particles(h))) return rc;
            sph_launch_density(h);
            if ((rc = non_pressure_acceleration(h, st))) return rc;
            sph_launch_update_velocity(h);
            sph_launch_wcsph_pressure(h);
            sph_launch_pressure_accel(h);
            sph_launch_update_velocity(h);
            sph_launch_update_position(h);
            if (h->c.has_dynamic_rigid) sph_launch_renew_rigid(h);
            sph_launch_boundary(h, SPH_MATERIAL_FLUID);
            break;
        case SPH_METHOD_PCISPH:  // PCISPH.py:165-185
            if ((rc = sph_sort_particles(h))) return rc;
            sph_launch_density(h);
            if ((rc = non_pressure_acceleration(h, st))) return rc;
            sph_launch_pcisph_init_step(h);
            if ((rc = pcisph_refine(h, &it, &e))) return rc;
            st->pcisph_iterations = it; st->pcisph_density_error = e; st->total_pcisph_iterations += it;
            sph_launch_update_velocity(h);
            sph_launch_pressure_accel(h);
            sph_launch_update_velocity(h);
            sph_launch_update_position(h);
            if (h->c.has_dynamic_rigid) sph_launch_renew_rigid(h);
            sph_launch_boundary(h, SPH_MATERIAL_FLUID);
            break;
        case SPH_METHOD_DFSPH:   // DFSPH.py:298-319
            if ((rc = non_pressure_acceleration(h, st))) return rc;
            sph_launch_update_velocity(h);
            if ((rc = dfsph_correct_density_error(h, &it, &e))) return rc;
            st->dfsph_iterations = it; st->dfsph_density_error = e; st->total_dfsph_iterations += it;
            sph_launch_update_position(h);
            if (h->c.has_dynamic_rigid) sph_launch_renew_rigid(h);
            sph_launch_boundary(h, SPH_MATERIAL_FLUID);
            if ((rc = sph_sort_particles(h))) return rc;
            sph_launch_density(h, true);   // + compute_alpha on the same staged windows
            if ((rc = dfsph_correct_divergence_error(h, &it, &e))) return rc;
            st->dfsph_iterations_v = it; st->dfsph_divergence_error = e; st->total_dfsph_iterations_v += it;
            break;
        default:
            return fail(h, SPH_E_UNSUPPORTED, "unknown simulation method");
    }
    // BaseSolver.step tail (base_solver.py:696).  Akinci volumes depend only on same-object rigid
    // neighbours: for static boundaries they are bit-identical every step (stable sort keeps the
    // rigid particles' relative order), so the sweep runs again only when something could change them.
    if (!h->rigid_volume_clean) {
        sph_launch_rigid_volume(h);
        h->rigid_volume_clean = !h->c.has_dynamic_rigid;
    }
    return last_launch(h);
}

struct FieldInfo { int comps; bool ok; };

FieldInfo field_info(int field) {
    switch (field) {
        case SPH_F_OBJECT_ID: case SPH_F_REST_VOLUME: case SPH_F_MASS: case SPH_F_DENSITY: case SPH_F_PRESSURE:
        case SPH_F_MATERIAL: case SPH_F_IS_DYNAMIC: case SPH_F_GRID_ID: case SPH_F_UID: case SPH_F_DFSPH_ALPHA:
        case SPH_F_DFSPH_KAPPA: case SPH_F_DFSPH_KAPPA_V: case SPH_F_DENSITY_STAR: case SPH_F_DENSITY_DERIVATIVE:
        case SPH_F_NEIGHBOR_COUNT:
            return {1, true};
        case SPH_F_POSITION: case SPH_F_VELOCITY: case SPH_F_ACCELERATION: case SPH_F_COLOR: case SPH_F_ORIGINAL_POSITION:
        case SPH_F_CELL: case SPH_F_PRESSURE_ACCELERATION: case SPH_F_PREDICTED_VELOCITY: case SPH_F_PREDICTED_POSITION:
        case SPH_F_CG_P: case SPH_F_ORIGINAL_VELOCITY: case SPH_F_CG_AP: case SPH_F_CG_X: case SPH_F_CG_B: case SPH_F_CG_R:
            return {3, true};
        case SPH_F_CG_DIAG_INV:
            return {9, true};
        default:
            return {0, false};
    }
}

}  // namespace

int sph_prof_begin(SphHandle* h, const char* name) {
    SphHandle::ProfRec r;
    r.name = name;
    for (cudaEvent_t* e : {&r.begin, &r.end}) {
        if (!h->event_pool.empty()) { *e = h->event_pool.back(); h->event_pool.pop_back(); }
        else if (cudaEventCreate(e) != cudaSuccess) return -1;
    }
    cudaEventRecord(r.begin, h->stream);
    h->prof.push_back(r);
    return (int)h->prof.size() - 1;
}
void sph_prof_end(SphHandle* h, int idx) { cudaEventRecord(h->prof[idx].end, h->stream); }

extern "C" {

int sph_abi_version(void) { return SPH_ABI_VERSION; }
const char* sph_backend_name(void) { return "cuda-sm100a"; }
const char* sph_last_error(const SphHandle* h) { return h ? h->err.c_str() : "null handle"; }

int sph_create(const SphParams* p, SphHandle** out) {
    if (!p || !out) return SPH_E_INVALID;
    if (p->abi_version != SPH_ABI_VERSION || p->dim != 3 || p->max_particles < 0) return SPH_E_INVALID;
    if (p->grid_num[0] <= 0 || p->grid_num[1] <= 0 || p->grid_num[2] <= 0) return SPH_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return SPH_E_CUDA;   // no CUDA device: this library has no CPU path
    }
    if (cudaSetDevice(p->device) != cudaSuccess) return SPH_E_CUDA;
    SphHandle* h = new SphHandle();
    h->P = *p;
    memset(&h->c, 0, sizeof h->c);
    memset(&h->d, 0, sizeof h->d);
    Consts& c = h->c;
    Dev& d = h->d;
    c.cap = p->max_particles;
    c.nx = p->grid_num[0]; c.ny = p->grid_num[1]; c.nz = p->grid_num[2];
    const long long ncell = (long long)c.nx * c.ny * c.nz;
    if (ncell > 2000000000LL) { delete h; return SPH_E_INVALID; }
    c.ncell = (int)ncell;
    c.z_lo = 0; c.z_hi = c.nz;
    refresh_consts(h);
    int rc = SPH_OK;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return SPH_E_CUDA; }
    h->own_stream = h->stream;
    const size_t n = (size_t)c.cap;
#define ALLOC(ptr, count) if (!rc) rc = dev_alloc(h, ptr, count)
    ALLOC(d.pv, n); ALLOC(d.pv_alt, n); ALLOC(d.vm, n); ALLOC(d.vm_alt, n);
    ALLOC(d.x0, 3 * n); ALLOC(d.x0_alt, 3 * n); ALLOC(d.rho, n); ALLOC(d.rho_alt, n);
    ALLOC(d.object_id, n); ALLOC(d.object_id_alt, n); ALLOC(d.material, n); ALLOC(d.material_alt, n);
    ALLOC(d.color, 3 * n); ALLOC(d.color_alt, 3 * n); ALLOC(d.is_dynamic, n); ALLOC(d.is_dynamic_alt, n);
    ALLOC(d.grid_id, n); ALLOC(d.grid_id_alt, n); ALLOC(d.uid, n); ALLOC(d.uid_alt, n);
    ALLOC(d.ghost_slot, n); ALLOC(d.ghost_slot_alt, n);
    ALLOC(d.acc, n); ALLOC(d.p, n);
    ALLOC(d.alpha, n); ALLOC(d.kappa, n); ALLOC(d.kappa_v, n); ALLOC(d.rho_star, n); ALLOC(d.drho, n);
    ALLOC(d.a_p, n); ALLOC(d.v_pred, n); ALLOC(d.x_pred, n);
    if (p->visc_method == SPH_VISC_IMPLICIT) {
        ALLOC(d.cg_p, n); ALLOC(d.v_orig, n); ALLOC(d.cg_Ap, n); ALLOC(d.cg_x, n); ALLOC(d.cg_b, n); ALLOC(d.cg_r, n);
        ALLOC(d.cg_dinv, 9 * n);
    }
    ALLOC(d.cell_count, (size_t)c.ncell + 8); ALLOC(d.cell_start, (size_t)c.ncell + 8);   // + the slab trash cell
    ALLOC(d.rank, n); ALLOC(d.perm, n);
    {
        const size_t m = (size_t)(c.ncell > c.cap ? c.ncell : c.cap);
        ALLOC(d.scan_tmp, m / 2048 + 1024);
    }
    ALLOC(d.object_material, SPH_MAX_OBJECTS); ALLOC(d.rigid_is_dynamic, SPH_MAX_OBJECTS);
    ALLOC(d.rigid_state, SPH_MAX_OBJECTS * 24); ALLOC(d.rigid_wrench, SPH_MAX_OBJECTS * 6);
    ALLOC(d.red, RED_COUNT);
    {   // neighbour lists: ELL, nbr_kmax slots x stride (SPH_B200_KMAX / SPH_B200_NO_LISTS override)
        const char* e = getenv("SPH_B200_NO_LISTS");
        h->lists_enabled = !(e && e[0] == '1');
        const char* k = getenv("SPH_B200_KMAX");
        d.nbr_kmax = k ? atoi(k) : 64;
        d.nbr_kmax = (d.nbr_kmax + 15) / 16 * 16;     // rows are read 16 slots (one 256-bit load) at a time
        if (d.nbr_kmax < 16) d.nbr_kmax = 16;
        if (h->lists_enabled) { ALLOC(d.nbr16, (size_t)d.nbr_kmax * n + 64); }
        ALLOC(d.aux, n);
        c.nbx = (c.nx + BRK_X - 1) / BRK_X; c.nby = (c.ny + BRK_Y - 1) / BRK_Y; c.nbz = (c.nz + BRK_Z - 1) / BRK_Z;
        h->nbricks = c.nbx * c.nby * c.nbz;
        ALLOC(d.brick_flag, (size_t)h->nbricks);
        ALLOC(d.brick_list, (size_t)h->nbricks);
        ALLOC(d.brick_ctl, BCTL_COUNT);
        ALLOC(d.brick_nf, (size_t)h->nbricks);
        if (h->lists_enabled) { ALLOC(d.row_order, n + 64); }
        const char* w = getenv("SPH_B200_WMAX");   // window slots per brick (16 B x staged arrays each)
        h->wmax = w ? atoi(w) : 2112;                 // 4 x 4 x 3 cells + halo = 180 cells x 10 particles at rest density + 17 %
        if (h->wmax < 64) h->wmax = 64;
        if (h->wmax > 4608) h->wmax = 4608;           // 3 arrays x 4608 x 16 B = 216 KB
        const char* bi = getenv("SPH_B200_BATCH_ITERS");
        h->solve_batch = bi ? atoi(bi) : 0;
    }
    if (!rc) {
        h->staging_bytes = (n ? n : 1) * 36;
        float* st = nullptr;
        rc = dev_alloc(h, st, h->staging_bytes / 4);
        h->staging = st;
    }
#undef ALLOC
    if (!rc && cudaMallocHost((void**)&h->h_red, sizeof(double) * RED_COUNT) != cudaSuccess) rc = SPH_E_CUDA;
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = SPH_E_CUDA;
    if (rc) {
        sph_destroy(h);
        return rc;
    }
    *out = h;
    return SPH_OK;
}

int sph_destroy(SphHandle* h) {
    if (!h) return SPH_OK;
    cudaSetDevice(h->P.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    sph_slab_free(h);
    for (void* p : h->allocations) cudaFree(p);
    if (h->h_red) cudaFreeHost(h->h_red);
    for (auto& r : h->prof) { cudaEventDestroy(r.begin); cudaEventDestroy(r.end); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return SPH_OK;
}

int sph_synchronize(SphHandle* h) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_add_particles(SphHandle* h, int32_t object_id, int32_t n, const float* x, const float* v, const float* density,
                      const float* pressure, const int32_t* material, const int32_t* is_dynamic, const int32_t* color) {
    if (!h || n < 0) return SPH_E_INVALID;
    ON_DEVICE(h);
    if (n == 0) return SPH_OK;
    if (!x || !v || !density || !pressure || !material || !is_dynamic || !color) return fail(h, SPH_E_INVALID, "null array");
    Consts& c = h->c;
    Dev& d = h->d;
    if ((long long)c.N + n > c.cap) return fail(h, SPH_E_CAPACITY, "particle_max_num exceeded");
    // add_particle (base_container.py:403-415): pack on the host, one async copy per array
    std::vector<float4> pv(n), vm(n);
    std::vector<int32_t> ids(n), uid(n);
    for (int k = 0; k < n; k++) {
        const float V = c.V0;
        pv[k] = make_float4(x[3 * k], x[3 * k + 1], x[3 * k + 2], material[k] == SPH_MATERIAL_FLUID ? V : -V);
        vm[k] = make_float4(v[3 * k], v[3 * k + 1], v[3 * k + 2], c.V0 * density[k]);
        ids[k] = object_id;
        uid[k] = c.N + k;
    }
    const size_t o = (size_t)c.N;
    cudaStream_t st = h->stream;
    CUDA_TRY(h, cudaMemcpyAsync(d.pv + o, pv.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.vm + o, vm.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.x0 + 3 * o, x, 12 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.rho + o, density, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.p + o, pressure, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.object_id + o, ids.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.uid + o, uid.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.material + o, material, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.is_dynamic + o, is_dynamic, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(d.color + 3 * o, color, 12 * (size_t)n, cudaMemcpyHostToDevice, st));
    // Z-slabs: the slots behind the live particles still carry the marks of what the last sort retired there
    CUDA_TRY(h, cudaMemsetAsync(d.ghost_slot + o, 0, 4 * (size_t)n, st));   // 0 = owned
    CUDA_TRY(h, cudaStreamSynchronize(st));   // host vectors go out of scope
    c.N += n;
    h->sorted_valid = false;
    h->list_valid = false;
    h->rigid_volume_clean = false;
    h->dyn_rigid_dirty = true;
    return SPH_OK;
}

int sph_get_field(SphHandle* h, int32_t field, void* dst, size_t bytes) {
    if (!h || !dst) return SPH_E_INVALID;
    ON_DEVICE(h);
    FieldInfo fi = field_info(field);
    if (!fi.ok) return fail(h, SPH_E_INVALID, "unknown field");
    const size_t item = (size_t)fi.comps * 4;
    if (bytes % item) return fail(h, SPH_E_INVALID, "byte size is not a whole number of items");
    const size_t n = bytes / item;
    if (n > (size_t)h->c.cap) return fail(h, SPH_E_INVALID, "size exceeds particle_max_num");
    if (n == 0) return SPH_OK;
    if (field == SPH_F_CELL) {
        sph_launch_cell_coords(h, (int*)h->staging, (int)n);
    } else if (field == SPH_F_NEIGHBOR_COUNT) {
        if (n > (size_t)h->c.N) return fail(h, SPH_E_INVALID, "neighbour counts exist for live particles only");
        sph_launch_neighbor_count(h, (int*)h->staging);
    } else {
        int comps = sph_field_to_staging(h, field, (int)n);
        if (comps < 0) return fail(h, comps, "field not available for this solver");
    }
    int rc = last_launch(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(dst, h->staging, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_set_field(SphHandle* h, int32_t field, const void* src, size_t bytes) {
    if (!h || !src) return SPH_E_INVALID;
    ON_DEVICE(h);
    FieldInfo fi = field_info(field);
    if (!fi.ok || field == SPH_F_CELL || field == SPH_F_NEIGHBOR_COUNT) return fail(h, SPH_E_INVALID, "field not settable");
    const size_t item = (size_t)fi.comps * 4;
    if (bytes % item) return fail(h, SPH_E_INVALID, "byte size is not a whole number of items");
    const size_t n = bytes / item;
    if (n > (size_t)h->c.cap) return fail(h, SPH_E_INVALID, "size exceeds particle_max_num");
    if (n == 0) return SPH_OK;
    CUDA_TRY(h, cudaMemcpyAsync(h->staging, src, bytes, cudaMemcpyHostToDevice, h->stream));
    int comps = sph_staging_to_field(h, field, (int)n);
    if (comps < 0) return fail(h, comps, "field not available for this solver");
    int rc = last_launch(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (field == SPH_F_POSITION) h->sorted_valid = false;
    if (field == SPH_F_POSITION || field == SPH_F_MATERIAL) h->list_valid = false;
    h->rigid_volume_clean = false;   // any host edit may touch what the boundary volumes depend on
    if (field == SPH_F_MATERIAL || field == SPH_F_IS_DYNAMIC) h->dyn_rigid_dirty = true;
    if (field == SPH_F_MATERIAL || field == SPH_F_REST_VOLUME || field == SPH_F_OBJECT_ID) h->bricks_dirty = true;
    return SPH_OK;
}

int sph_fill_field(SphHandle* h, int32_t field, double value) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    FieldInfo fi = field_info(field);
    if (!fi.ok || field == SPH_F_CELL || field == SPH_F_NEIGHBOR_COUNT) return fail(h, SPH_E_INVALID, "field not settable");
    const size_t n = (size_t)h->c.cap;
    const bool is_int = field == SPH_F_OBJECT_ID || field == SPH_F_MATERIAL || field == SPH_F_COLOR || field == SPH_F_IS_DYNAMIC ||
                        field == SPH_F_GRID_ID || field == SPH_F_UID;
    if (n == 0) return SPH_OK;
    if (is_int) sph_fill_i32(h, (int*)h->staging, n * fi.comps, (int)value);
    else sph_fill_f32(h, (float*)h->staging, n * fi.comps, (float)value);
    int comps = sph_staging_to_field(h, field, (int)n);
    if (comps < 0) return fail(h, comps, "field not available for this solver");
    if (field == SPH_F_MATERIAL || field == SPH_F_IS_DYNAMIC) h->dyn_rigid_dirty = true;
    if (field == SPH_F_MATERIAL || field == SPH_F_REST_VOLUME || field == SPH_F_OBJECT_ID) h->bricks_dirty = true;
    if (field == SPH_F_POSITION || field == SPH_F_MATERIAL) h->list_valid = false;
    h->rigid_volume_clean = false;   // any host edit may touch what the boundary volumes depend on
    return last_launch(h);
}

int sph_field_ptr(SphHandle* h, int32_t field, void** ptr, int32_t* stride, int32_t* comps) {
    if (!h) return SPH_E_INVALID;
    Dev& d = h->d;
    void* p = nullptr;
    int s = 4, cmp = 1;
    switch (field) {
        case SPH_F_POSITION: p = d.pv; s = 16; cmp = 3; break;
        case SPH_F_VELOCITY: p = d.vm; s = 16; cmp = 3; break;
        case SPH_F_ACCELERATION: p = d.acc; s = 16; cmp = 3; break;
        case SPH_F_DENSITY: p = d.rho; break;
        case SPH_F_PRESSURE: p = d.p; break;
        case SPH_F_MATERIAL: p = d.material; break;
        case SPH_F_OBJECT_ID: p = d.object_id; break;
        case SPH_F_UID: p = d.uid; break;
        case SPH_F_GRID_ID: p = d.grid_id; break;
        case SPH_F_IS_DYNAMIC: p = d.is_dynamic; break;
        case SPH_F_DFSPH_ALPHA: p = d.alpha; break;
        case SPH_F_DENSITY_STAR: p = d.rho_star; break;
        default: return fail(h, SPH_E_INVALID, "no raw pointer for this field");
    }
    if (ptr) *ptr = p;
    if (stride) *stride = s;
    if (comps) *comps = cmp;
    return SPH_OK;
}

int sph_get_scalar(SphHandle* h, int32_t s, double* out) {
    if (!h || !out) return SPH_E_INVALID;
    ON_DEVICE(h);
    int rc;
    switch (s) {
        case SPH_S_DT: *out = h->c.dt; break;
        case SPH_S_PARTICLE_NUM: *out = h->c.N; break;
        case SPH_S_FLUID_PARTICLE_NUM: *out = h->Nfluid; break;
        case SPH_S_PCISPH_K: *out = h->c.pcisph_k; break;
        case SPH_S_DENSITY_ERROR: *out = h->density_error; break;
        case SPH_S_CG_ALPHA: {
            if ((rc = read_red(h))) return rc;
            const float num = (float)h->h_red[RED_CG_RR], den = (float)h->h_red[RED_CG_PAP];
            *out = den > 1e-18f ? num / den : 0.0f;
            break;
        }
        case SPH_S_CG_BETA: {
            if ((rc = read_red(h))) return rc;
            const float num = (float)h->h_red[RED_CG_RR_NEW], den = (float)h->h_red[RED_CG_RR_OLD];
            *out = den > 1e-18f ? num / den : 0.0f;
            break;
        }
        case SPH_S_CG_ERROR: *out = h->cg_error; break;
        case SPH_S_G_UPPER: *out = h->c.g_upper; break;
        case SPH_S_VISCOSITY: *out = h->P.viscosity; break;
        case SPH_S_VISCOSITY_B: *out = h->P.viscosity_b; break;
        case SPH_S_NUM_CELLS: *out = h->c.ncell; break;
        case SPH_S_MAX_PARTICLES: *out = h->c.cap; break;
        case SPH_S_ACTIVE_BRICKS: case SPH_S_MAX_WINDOW_SLOTS: case SPH_S_WINDOW_OVERFLOWS: {
            int ctl[BCTL_COUNT];
            CUDA_TRY(h, cudaMemcpyAsync(ctl, h->d.brick_ctl, sizeof ctl, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            *out = ctl[s == SPH_S_ACTIVE_BRICKS ? BCTL_ACTIVE : (s == SPH_S_MAX_WINDOW_SLOTS ? BCTL_WMAX_SEEN : BCTL_OVERFLOWS)];
            break;
        }
        default: return fail(h, SPH_E_INVALID, "unknown scalar");
    }
    return SPH_OK;
}

int sph_set_scalar(SphHandle* h, int32_t s, double v) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    switch (s) {
        case SPH_S_DT: h->P.dt = v; refresh_consts(h); break;
        case SPH_S_PARTICLE_NUM:
            if (v < 0 || v > h->c.cap) return fail(h, SPH_E_CAPACITY, "particle_num out of range");
            h->c.N = (int)v; h->rigid_volume_clean = false; h->sorted_valid = false; h->list_valid = false; break;
        case SPH_S_FLUID_PARTICLE_NUM: h->Nfluid = (int)v; break;
        case SPH_S_PCISPH_K: h->c.pcisph_k = (float)v; break;
        case SPH_S_DENSITY_ERROR: h->density_error = (float)v; break;
        case SPH_S_CG_ERROR: h->cg_error = (float)v; break;
        case SPH_S_G_UPPER: h->P.g_upper = v; refresh_consts(h); break;
        case SPH_S_VISCOSITY: h->P.viscosity = v; refresh_consts(h); break;
        case SPH_S_VISCOSITY_B: h->P.viscosity_b = v; refresh_consts(h); break;
        default: return fail(h, SPH_E_INVALID, "scalar not settable");
    }
    return SPH_OK;
}

int sph_set_object(SphHandle* h, int32_t obj, int32_t material, int32_t is_dynamic) {
    if (!h || obj < 0 || obj >= SPH_MAX_OBJECTS) return SPH_E_INVALID;
    ON_DEVICE(h);
    h->object_material[obj] = material;
    h->rigid_is_dynamic_h[obj] = is_dynamic;
    h->bricks_dirty = true;   // emitter objects (fluid objects parked as rigid) count as working rows
    CUDA_TRY(h, cudaMemcpyAsync(h->d.object_material, h->object_material, sizeof h->object_material, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d.rigid_is_dynamic, h->rigid_is_dynamic_h, sizeof h->rigid_is_dynamic_h, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_set_rigid_state(SphHandle* h, int32_t obj, const float com0[3], const float com[3], const float rot[9],
                        const float vel[3], const float omega[3]) {
    if (!h || obj < 0 || obj >= SPH_MAX_OBJECTS) return SPH_E_INVALID;
    ON_DEVICE(h);
    float* s = h->rigid_state_h[obj];
    if (com0) memcpy(s + 0, com0, 12);
    if (com) memcpy(s + 3, com, 12);
    if (rot) memcpy(s + 6, rot, 36);
    if (vel) memcpy(s + 15, vel, 12);
    if (omega) memcpy(s + 18, omega, 12);
    CUDA_TRY(h, cudaMemcpyAsync(h->d.rigid_state + obj * 24, s, 24 * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_get_rigid_wrench(SphHandle* h, float* force, float* torque) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    float w[SPH_MAX_OBJECTS * 6];
    CUDA_TRY(h, cudaMemcpyAsync(w, h->d.rigid_wrench, sizeof w, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (int o = 0; o < SPH_MAX_OBJECTS; o++)
        for (int k = 0; k < 3; k++) {
            if (force) force[3 * o + k] = w[6 * o + k];
            if (torque) torque[3 * o + k] = w[6 * o + 3 + k];
        }
    return SPH_OK;
}

int sph_zero_rigid_wrench(SphHandle* h) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    CUDA_TRY(h, cudaMemsetAsync(h->d.rigid_wrench, 0, sizeof(float) * SPH_MAX_OBJECTS * 6, h->stream));
    return SPH_OK;
}

int sph_compute_rigid_body_mass(SphHandle* h, int32_t object_id, float* out) {
    if (!h || !out) return SPH_E_INVALID;
    ON_DEVICE(h);
    int rc;
    if ((rc = zero_red(h, RED_MASS))) return rc;
    sph_launch_rigid_body_mass(h, object_id);
    if ((rc = read_red(h))) return rc;
    *out = (float)h->h_red[RED_MASS];
    return SPH_OK;
}

int sph_prepare_neighborhood_search(SphHandle* h) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    int rc = update_dynamic_rigid_flag(h);
    if (rc) return rc;
    rc = sph_sort_particles(h);
    if (rc) return check_cuda(h, cudaGetLastError(), "sort");
    return last_launch(h);
}

int sph_get_neighbors(SphHandle* h, int32_t* offsets, int32_t* indices, size_t capacity) {
    if (!h || !offsets) return SPH_E_INVALID;
    ON_DEVICE(h);
    const int N = h->c.N;
    int *counts = nullptr, *offs = nullptr, *idx = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&counts, sizeof(int) * (size_t)(N + 1)));
    cudaError_t e = cudaMalloc((void**)&offs, sizeof(int) * (size_t)(N + 2));
    if (e != cudaSuccess) { cudaFree(counts); return check_cuda(h, e, "cudaMalloc"); }
    sph_launch_neighbor_count(h, counts);
    sph_exclusive_scan(h, counts, N, offs);
    int rc = check_cuda(h, cudaMemcpyAsync(offsets, offs, sizeof(int) * (size_t)(N + 1), cudaMemcpyDeviceToHost, h->stream), "copy of the neighbour offsets");
    if (!rc) rc = check_cuda(h, cudaStreamSynchronize(h->stream), "neighbour count");
    if (N == 0) offsets[0] = 0;
    if (!rc) rc = last_launch(h);
    if (!rc && indices) {
        const size_t total = (size_t)offsets[N];
        if (total > capacity) rc = fail(h, SPH_E_CAPACITY, "indices capacity too small");
        else if (total) {
            e = cudaMalloc((void**)&idx, sizeof(int) * total);
            if (e != cudaSuccess) rc = check_cuda(h, e, "cudaMalloc");
            else {
                sph_launch_neighbor_fill(h, offs, idx);
                rc = check_cuda(h, cudaMemcpyAsync(indices, idx, sizeof(int) * total, cudaMemcpyDeviceToHost, h->stream), "copy of the neighbour indices");
                if (!rc) rc = check_cuda(h, cudaStreamSynchronize(h->stream), "neighbour fill");
                if (!rc) rc = last_launch(h);
            }
        }
    }
    cudaFree(counts); cudaFree(offs);
    if (idx) cudaFree(idx);
    return rc;
}

int sph_get_grid_num_particles(SphHandle* h, int32_t* dst, size_t count) {
    if (!h || !dst || count > (size_t)h->c.ncell) return SPH_E_INVALID;
    ON_DEVICE(h);
    // per-cell histogram in the reference's z-fastest flatten, inclusive-scanned (base_container.py:546)
    int *hist = nullptr, *scan = nullptr;
    const size_t nc = (size_t)h->c.ncell;
    CUDA_TRY(h, cudaMalloc((void**)&hist, sizeof(int) * (nc + 8)));
    cudaError_t e = cudaMalloc((void**)&scan, sizeof(int) * (nc + 8));
    if (e != cudaSuccess) { cudaFree(hist); return check_cuda(h, e, "cudaMalloc"); }
    cudaMemsetAsync(hist, 0, sizeof(int) * (nc + 8), h->stream);
    sph_launch_ref_cell_hist(h, hist);
    sph_exclusive_scan(h, hist, (int)nc, scan);
    int rc = check_cuda(h, cudaMemcpyAsync(dst, scan + 1, sizeof(int) * count, cudaMemcpyDeviceToHost, h->stream), "copy of the cell counts");   // inclusive = exclusive shifted by one
    if (!rc) rc = check_cuda(h, cudaStreamSynchronize(h->stream), "cell histogram");
    if (!rc) rc = last_launch(h);
    cudaFree(hist); cudaFree(scan);
    return rc;
}

int sph_run_task(SphHandle* h, int32_t task, int32_t iarg, float* out) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    if (h->slab && ((task >= SPH_T_CG_PREPARE1 && task <= SPH_T_COPY_BACK_ORIGINAL_VELOCITY) || task >= SPH_T_PCISPH_COMPUTE_PREDICTED_VELOCITY))
        return fail(h, SPH_E_UNSUPPORTED, "Z-slab handles support WCSPH and DFSPH with standard viscosity");
    int rc = update_dynamic_rigid_flag(h);
    if (rc) return rc;
    const bool implicit = h->P.visc_method == SPH_VISC_IMPLICIT;
    switch (task) {   // neighbour sweeps read the chunk windows laid out by the last sort
        case SPH_T_COMPUTE_PRESSURE_ACCELERATION: case SPH_T_COMPUTE_SURFACE_TENSION_ACCELERATION:
        case SPH_T_COMPUTE_VISCOSITY_ACCELERATION_STANDARD: case SPH_T_COMPUTE_DENSITY: case SPH_T_CG_PREPARE1:
        case SPH_T_CG_COMPUTE_AP: case SPH_T_DFSPH_COMPUTE_ALPHA: case SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE:
        case SPH_T_DFSPH_COMPUTE_DENSITY_STAR: case SPH_T_DFSPH_CORRECT_DIVERGENCE_STEP:
        case SPH_T_DFSPH_CORRECT_DENSITY_ERROR_STEP: case SPH_T_PCISPH_COMPUTE_DENSITY_STAR:
        case SPH_T_PCISPH_COMPUTE_TEMP_PRESSURE_ACCELERATION:
            if (!h->sorted_valid)
                return fail(h, SPH_E_STATE, "particles were added or moved by the host since the last sort: call prepare_neighborhood_search() first");
            break;
        default: break;
    }
    if (task >= SPH_T_CG_PREPARE1 && task <= SPH_T_COPY_BACK_ORIGINAL_VELOCITY && !implicit)
        return fail(h, SPH_E_STATE, "implicit-viscosity kernels need viscosityMethod = implicit");
    switch (task) {
        case SPH_T_COMPUTE_RIGID_PARTICLE_VOLUME: sph_launch_rigid_volume(h); break;
        case SPH_T_COMPUTE_PRESSURE_ACCELERATION: sph_launch_pressure_accel(h); break;
        case SPH_T_COMPUTE_GRAVITY_ACCELERATION: sph_launch_gravity(h); break;
        case SPH_T_COMPUTE_SURFACE_TENSION_ACCELERATION: sph_launch_surface_tension(h); break;
        case SPH_T_COMPUTE_VISCOSITY_ACCELERATION_STANDARD: sph_launch_viscosity(h); break;
        case SPH_T_COMPUTE_DENSITY: sph_launch_density(h); break;
        case SPH_T_ENFORCE_DOMAIN_BOUNDARY_3D: sph_launch_boundary(h, iarg); break;
        case SPH_T_RENEW_RIGID_PARTICLE_STATE: sph_launch_renew_rigid(h); break;
        case SPH_T_UPDATE_FLUID_VELOCITY: sph_launch_update_velocity(h); break;
        case SPH_T_UPDATE_FLUID_POSITION: sph_launch_update_position(h); break;
        case SPH_T_PREPARE_EMITTER: sph_launch_prepare_emitter(h); h->dyn_rigid_dirty = true; h->rigid_volume_clean = false; break;
        case SPH_T_INIT_OBJECT_ID: sph_fill_i32(h, h->d.object_id, (size_t)h->c.cap, -1); h->rigid_volume_clean = false; break;
        case SPH_T_INIT_ACCELERATION: sph_fill_f32(h, (float*)h->d.acc, (size_t)h->c.cap * 4, 0.0f); break;
        case SPH_T_INIT_RIGID_BODY_FORCE_AND_TORQUE: return sph_zero_rigid_wrench(h);
        case SPH_T_CG_PREPARE1: sph_launch_cg_prepare1_pre(h); sph_launch_cg_prepare1(h); break;
        case SPH_T_CG_PREPARE2: sph_launch_cg_prepare2(h); break;
        case SPH_T_CG_COMPUTE_AP: sph_launch_cg_Ap(h, false); break;
        case SPH_T_CG_COMPUTE_ALPHA: if ((rc = cg_dots(h))) return rc; break;
        case SPH_T_CG_UPDATE_X: sph_launch_cg_update_x(h); break;
        case SPH_T_CG_UPDATE_R_AND_BETA: if ((rc = cg_update_r(h))) return rc; if (out) *out = h->cg_error; break;
        case SPH_T_CG_UPDATE_P: sph_launch_cg_update_p(h); break;
        case SPH_T_CG_PREPARE_GUESS: sph_launch_cg_prepare_guess(h); break;
        case SPH_T_VISCOSITY_UPDATE_VELOCITY: sph_launch_cg_velocity_from_x(h); break;
        case SPH_T_COPY_BACK_ORIGINAL_VELOCITY: sph_launch_cg_velocity_restore(h); break;
        case SPH_T_WCSPH_COMPUTE_PRESSURE: sph_launch_wcsph_pressure(h); break;
        case SPH_T_DFSPH_COMPUTE_ALPHA: sph_launch_dfsph_alpha(h); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE: sph_launch_dfsph_density_derivative(h, false); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_STAR: sph_launch_dfsph_density_star(h, false); break;
        case SPH_T_DFSPH_COMPUTE_KAPPA_V: sph_launch_dfsph_kappa_v(h); break;
        case SPH_T_DFSPH_CORRECT_DIVERGENCE_STEP: sph_launch_dfsph_correct_divergence(h, false); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE_ERROR:
        case SPH_T_DFSPH_COMPUTE_DENSITY_ERROR:
            if ((rc = zero_red(h, RED_ERR))) return rc;
            if (task == SPH_T_DFSPH_COMPUTE_DENSITY_ERROR) sph_launch_dfsph_density_error(h);
            else sph_launch_dfsph_divergence_error(h);
            if ((rc = sph_slab_allreduce_red(h, RED_ERR, 1))) return rc;
            if ((rc = read_red(h))) return rc;
            if (out) *out = (float)h->h_red[RED_ERR] / global_particle_num(h);
            break;
        case SPH_T_DFSPH_COMPUTE_KAPPA: sph_launch_dfsph_kappa(h); break;
        case SPH_T_DFSPH_CORRECT_DENSITY_ERROR_STEP: sph_launch_dfsph_correct_density(h, false); break;
        case SPH_T_PCISPH_COMPUTE_PREDICTED_VELOCITY: sph_launch_pcisph_predict_velocity(h); break;
        case SPH_T_PCISPH_COMPUTE_PREDICTED_POSITION: sph_launch_pcisph_predict_position(h); break;
        case SPH_T_PCISPH_COMPUTE_DENSITY_STAR:
            if ((rc = pcisph_density_star(h))) return rc;
            if ((rc = pcisph_fetch_error(h))) return rc;
            if (out) *out = h->density_error;
            break;
        case SPH_T_PCISPH_UPDATE_PRESSURE: sph_launch_pcisph_update_pressure(h); break;
        case SPH_T_PCISPH_COMPUTE_TEMP_PRESSURE_ACCELERATION: sph_launch_temp_pressure_accel(h); break;
        case SPH_T_PCISPH_COMPUTE_K: h->c.pcisph_k = pcisph_k_host(h); if (out) *out = h->c.pcisph_k; break;
        case SPH_T_PCISPH_INIT_STEP: sph_launch_pcisph_init_step(h); h->density_error = 100.0f; break;
        default: return fail(h, SPH_E_INVALID, "unknown task");
    }
    return last_launch(h);
}

int sph_step(SphHandle* h, int32_t n_steps, SphStepStats* stats) {
    if (!h || n_steps < 0) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    if (h->slab && (h->P.method == SPH_METHOD_PCISPH || h->P.visc_method == SPH_VISC_IMPLICIT))
        return fail(h, SPH_E_UNSUPPORTED, "Z-slab handles support WCSPH and DFSPH with standard viscosity");
    // a DFSPH step begins with sweeps on the grid of the previous step's sort (DFSPH.py:298-319)
    if (h->P.method == SPH_METHOD_DFSPH && !h->sorted_valid && n_steps > 0)
        return fail(h, SPH_E_STATE, "particles were added or moved by the host since the last sort: call prepare_neighborhood_search() first");
    int rc = update_dynamic_rigid_flag(h);
    if (rc) return rc;
    SphStepStats st;
    memset(&st, 0, sizeof st);
    const int64_t l0 = h->launches;
    for (int k = 0; k < n_steps; k++) {
        if ((rc = step_once(h, &st))) return rc;
        st.steps++;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    st.kernel_launches = h->launches - l0;
    if (stats) *stats = st;
    return SPH_OK;
}

int sph_dfsph_correct_density_error(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    int i; float err;
    int rc = dfsph_correct_density_error(h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return rc;
}
int sph_dfsph_correct_divergence_error(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    int i; float err;
    int rc = dfsph_correct_divergence_error(h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return rc;
}
int sph_pcisph_refine(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    int i; float err;
    int rc = pcisph_refine(h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return rc;
}
int sph_implicit_viscosity_solve(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    rows_all(h);
    if (h->P.visc_method != SPH_VISC_IMPLICIT) return fail(h, SPH_E_STATE, "viscosityMethod is not implicit");
    int i; float err;
    int rc = implicit_viscosity_solve(h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return rc;
}

int sph_set_stream(SphHandle* h, void* cuda_stream) {
    if (!h) return SPH_E_INVALID;
    ON_DEVICE(h);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return SPH_OK;
}

int sph_profile_enable(SphHandle* h, int32_t enable) {
    if (!h) return SPH_E_INVALID;
    h->profiling = enable != 0;
    return SPH_OK;
}

int sph_profile_read(SphHandle* h, SphKernelStat* out, int32_t capacity, int32_t* count) {
    if (!h || !count) return SPH_E_INVALID;
    ON_DEVICE(h);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int n = 0;
    for (auto& r : h->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.begin, r.end) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
        char name[sizeof out[0].name];   // macro arguments may arrive parenthesised: "(kb_build<true, true, false>)"
        memset(name, 0, sizeof name);
        strncpy(name, r.name + (r.name[0] == '(' ? 1 : 0), sizeof name - 1);
        if (char* paren = strchr(name, ')')) *paren = 0;
        int k = 0;
        for (; k < n; k++) if (!strncmp(out[k].name, name, sizeof name - 1)) break;
        if (k == n) {
            if (n >= capacity || !out) { h->event_pool.push_back(r.begin); h->event_pool.push_back(r.end); continue; }
            memset(&out[n], 0, sizeof out[n]);
            memcpy(out[n].name, name, sizeof name);
            n++;
        }
        out[k].launches++;
        out[k].total_ms += ms;
        h->event_pool.push_back(r.begin);
        h->event_pool.push_back(r.end);
    }
    h->prof.clear();
    *count = n;
    return SPH_OK;
}

// ---- Z-slab sharding: implemented in sph_slab.cu ------------------------------------------------

}  // extern "C"
