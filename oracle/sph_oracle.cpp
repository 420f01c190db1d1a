// sph_oracle.cpp — CPU restatement of the reference's SPH hot path.  TEST INFRASTRUCTURE ONLY.
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN PYTHON SOURCES, with one caveat.  The reference
// (jason-huang03/SPH_Project @ 2a97e63) ships no tests, golden vectors or fixtures, and Taichi /
// pybullet / trimesh cannot be installed in this image, so the real Taichi runtime never ran here.
// Instead tests/golden/make_ref_golden.py imports /root/reference/SPH in place and steps it on
// tests/golden/ref_shim, a small emulation of the Taichi API (kernel bodies executed as serial Python
// on f32 numpy scalars).  The committed fixtures tests/golden/ref_*.npz hold its state after prepare()
// and after every step for fifteen tiny WCSPH / PCISPH / DFSPH scenes and BASELINE config C1 (domain box, over-dense blocks that
// make the pressure solvers iterate, implicit viscosity, emitter, late-entry block, mesh bodies, a dynamic
// rigid cube coupled to each solver); tests/test_ref_golden.py checks this
// file against them: insertion lattice and integer fields bit for bit, every float field within 2e-5
// of its scale (positions 2e-7), solver iteration counts exactly.  The caveat: Taichi's code
// generation (fast-math, fma contraction, pow / inverse lowering, atomic order) is emulated, not run.
// Further pins: (i) closed-form known answers (tests/test_oracle_kat.py), (ii) an independent O(N^2)
// numpy restatement of the same formulas (oracle/bruteforce.py), (iii) the scene-count known answers
// of SURVEY.md 8(c).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (sph_project_b200) never does.
//
// Arithmetic: f32 like the reference (ti.init default_fp = f32), evaluated left to right as the
// Python expressions are written, Python-side constants computed in f64 and rounded to f32 where
// they meet an f32 operand.  Build with -ffp-contract=off so the only fused operation is the
// explicit fmaf in dist2() (the canonical squared distance shared with the CUDA path so both
// produce bit-identical neighbour sets).  Reductions that the reference does with f32 atomics in
// nondeterministic order are accumulated in f64 (OpenMP partial sums) and rounded once.
//
// Sort order: stable counting sort on the reference's z-fastest flatten (base_container.py:472-481);
// upstream's in-cell order is nondeterministic (atomic_sub from many threads, :510-515) but equals
// this stable order when its loop runs serially.
//
// Paths cited below are relative to the reference checkout.

#include "../include/sph_b200.h"

#include <algorithm>
#include <parallel/algorithm>
#include <type_traits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include <sys/syscall.h>
#include <unistd.h>

namespace {

struct V3 {
    float x, y, z;
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator-=(V3& a, V3 b) { a = a - b; return a; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// canonical squared distance (shared with the CUDA path): fma chain in x, y, z order
static inline float dist2(V3 r) { return fmaf(r.z, r.z, fmaf(r.y, r.y, r.x * r.x)); }
static inline float norm(V3 r) { return sqrtf(dist2(r)); }
static inline float norm_sqr(V3 r) { return dist2(r); }

struct M3 {
    float m[9];
};
static inline M3 m3_zero() { M3 r; for (float& v : r.m) v = 0.f; return r; }
static inline M3 m3_identity() { M3 r = m3_zero(); r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
static inline M3 outer(V3 a, V3 b) {
    return {{a.x * b.x, a.x * b.y, a.x * b.z, a.y * b.x, a.y * b.y, a.y * b.z, a.z * b.x, a.z * b.y, a.z * b.z}};
}
static inline M3 operator*(M3 a, float s) { for (float& v : a.m) v *= s; return a; }
static inline M3 operator/(M3 a, float s) { for (float& v : a.m) v /= s; return a; }
static inline M3 operator-(M3 a, M3 b) { for (int i = 0; i < 9; i++) a.m[i] -= b.m[i]; return a; }
static inline M3 operator-(M3 a) { for (float& v : a.m) v = -v; return a; }
static inline V3 mv(const M3& a, V3 v) {
    return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
static inline M3 mm(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r.m[i * 3 + j] = a.m[i * 3 + 0] * b.m[0 * 3 + j] + a.m[i * 3 + 1] * b.m[1 * 3 + j] + a.m[i * 3 + 2] * b.m[2 * 3 + j];
    return r;
}
// ti.math.inverse for 3x3 (base_solver.py:308): adjugate / determinant
static inline M3 inverse(const M3& a) {
    const float* m = a.m;
    float c00 = m[4] * m[8] - m[5] * m[7];
    float c01 = m[5] * m[6] - m[3] * m[8];
    float c02 = m[3] * m[7] - m[4] * m[6];
    float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    float inv = 1.0f / det;
    M3 r;
    r.m[0] = c00 * inv;
    r.m[1] = (m[2] * m[7] - m[1] * m[8]) * inv;
    r.m[2] = (m[1] * m[5] - m[2] * m[4]) * inv;
    r.m[3] = c01 * inv;
    r.m[4] = (m[0] * m[8] - m[2] * m[6]) * inv;
    r.m[5] = (m[2] * m[3] - m[0] * m[5]) * inv;
    r.m[6] = c02 * inv;
    r.m[7] = (m[1] * m[6] - m[0] * m[7]) * inv;
    r.m[8] = (m[0] * m[4] - m[1] * m[3]) * inv;
    return r;
}

}  // namespace

struct SphHandle {
    SphParams P;
    std::string err;
    int N = 0;        // particle_num[None]
    int Nfluid = 0;   // fluid_particle_num[None]
    int cap = 0;
    int ncell = 0;
    // constants rounded to f32 once
    float h, dx, diameter, V0, rho0, dt, g_upper, visc, visc_b, sigma, padding;
    V3 g, dom;
    float kW, kW2, kG;  // 8/(pi h^3), 2*that, 6*that
    // particle fields (base_container.py:138-148,155,183)
    std::vector<int32_t> object_id, material, is_dynamic, grid_id, uid, color;  // color: 3 per particle
    std::vector<V3> x, v, a, x0;
    std::vector<float> V, m, rho, p;
    // dfsph_container.py:13-17 / pcisph_container.py:14-19
    std::vector<float> alpha, kappa, kappa_v, rho_star, drho;
    std::vector<V3> a_p, v_pred, x_pred;
    float pcisph_k = 0.f, density_error = 0.f;
    // base_solver.py:43-52
    std::vector<V3> cg_p, v_orig, cg_Ap, cg_x, cg_b, cg_r;
    std::vector<M3> cg_dinv;
    float cg_alpha = 0.f, cg_beta = 0.f, cg_error = 0.f;
    // grid (base_container.py:132-133): inclusive scan of per-cell counts, z-fastest flatten
    std::vector<int32_t> cell_scan;
    // object tables (base_container.py:150-165)
    int32_t object_material[SPH_MAX_OBJECTS] = {0};
    int32_t rigid_is_dynamic[SPH_MAX_OBJECTS] = {0};
    V3 rigid_com0[SPH_MAX_OBJECTS], rigid_com[SPH_MAX_OBJECTS], rigid_vel[SPH_MAX_OBJECTS], rigid_omega[SPH_MAX_OBJECTS];
    M3 rigid_rot[SPH_MAX_OBJECTS];
    double rigid_force[SPH_MAX_OBJECTS][3], rigid_torque[SPH_MAX_OBJECTS][3];

    // ---- kernel functions (base_solver.py:56-103) ----
    float kernel_W(float R_mod) const {
        float res = 0.f;
        float q = R_mod / h;
        if (q <= 1.0f) {
            if (q <= 0.5f) {
                float q2 = q * q;
                float q3 = q2 * q;
                res = kW * (6.0f * q3 - 6.0f * q2 + 1.f);
            } else {
                res = kW2 * powf(1.f - q, 3.0f);
            }
        }
        return res;
    }
    V3 kernel_gradient(V3 R) const {
        float R_mod = norm(R);
        float q = R_mod / h;
        V3 res = {0.f, 0.f, 0.f};
        if (R_mod > 1e-5f && q <= 1.0f) {
            V3 grad_q = R / (R_mod * h);
            if (q <= 0.5f) {
                res = (kG * q * (3.0f * q - 2.0f)) * grad_q;
            } else {
                float factor = 1.0f - q;
                res = (kG * (-factor * factor)) * grad_q;
            }
        }
        return res;
    }

    // ---- grid (base_container.py:467-481) ----
    void pos_to_index(V3 pos, int c[3]) const {
        c[0] = (int)(pos.x / h);  // .cast(int) truncates toward zero
        c[1] = (int)(pos.y / h);
        c[2] = (int)(pos.z / h);
    }
    int flatten(const int c[3]) const { return (c[0] * P.grid_num[1] + c[1]) * P.grid_num[2] + c[2]; }

    // for_all_neighbors (base_container.py:549-560).  Out-of-range neighbour cells are skipped
    // (upstream does no bounds check: SURVEY.md App. B#5; identical wherever upstream is defined).
    template <class F>
    void for_all_neighbors(int p_i, F&& task) const {
        int c[3];
        pos_to_index(x[p_i], c);
        for (int ox = -1; ox <= 1; ox++)
            for (int oy = -1; oy <= 1; oy++)
                for (int oz = -1; oz <= 1; oz++) {
                    int n[3] = {c[0] + ox, c[1] + oy, c[2] + oz};
                    if (n[0] < 0 || n[1] < 0 || n[2] < 0 || n[0] >= P.grid_num[0] || n[1] >= P.grid_num[1] || n[2] >= P.grid_num[2]) continue;
                    int gi = flatten(n);
                    int start = gi - 1 >= 0 ? cell_scan[gi - 1] : 0;
                    int end = cell_scan[gi];
                    for (int p_j = start; p_j < end; p_j++)
                        if (p_i != p_j && norm(x[p_i] - x[p_j]) < h) task(p_j);
                }
    }
};

namespace {

int fail(SphHandle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}

// init_grid + PrefixSumExecutor.run + reorder_particles (base_container.py:495-547)
//
// The reference's reorder is a counting sort whose in-cell order, run serially, is ascending particle
// index (descending loop + atomic_sub, :510-515).  The same permutation is obtained here by sorting the
// unique 64-bit keys (cell << 32 | index), which parallelises (libstdc++ parallel mode) -- the serial
// counting sort was the Amdahl bottleneck of the CPU baseline on many-core hosts.
void prepare_neighborhood_search(SphHandle& s) {
    const int N = s.N;
    std::vector<uint64_t> keys((size_t)N);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        int c[3];
        s.pos_to_index(s.x[i], c);
        // keep indices inside the array like a well-defined run of the reference would
        for (int d = 0; d < 3; d++) c[d] = std::min(std::max(c[d], 0), s.P.grid_num[d] - 1);
        s.grid_id[i] = s.flatten(c);
        keys[i] = ((uint64_t)(uint32_t)s.grid_id[i] << 32) | (uint32_t)i;
    }
    __gnu_parallel::sort(keys.begin(), keys.end());
    // per-cell counts -> inclusive scan (:546): cell c ends where the last key of a cell <= c sits
#pragma omp parallel for schedule(static)
    for (int c = 0; c < s.ncell; c++) s.cell_scan[c] = 0;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < N; k++) {
        int cell = (int)(keys[k] >> 32);
        if (k + 1 == N || (int)(keys[k + 1] >> 32) != cell) s.cell_scan[cell] = k + 1;
    }
    int acc = 0;     // empty cells inherit the running end
    for (int c = 0; c < s.ncell; c++) {
        if (s.cell_scan[c] == 0) s.cell_scan[c] = acc;
        acc = s.cell_scan[c];
    }
    // permute exactly the fields of :517-542 (+ uid); everything else stays index-bound
    auto permute = [&](auto& vec, int comps) {
        using T = typename std::remove_reference<decltype(vec)>::type::value_type;
        std::vector<T> buf(vec.size());
#pragma omp parallel for schedule(static)
        for (int k = 0; k < N; k++) {
            size_t src = (size_t)(uint32_t)keys[k];
            for (int c = 0; c < comps; c++) buf[(size_t)k * comps + c] = vec[src * comps + c];
        }
        size_t live = (size_t)N * comps;
#pragma omp parallel for schedule(static)
        for (long long q = (long long)live; q < (long long)vec.size(); q++) buf[q] = vec[q];
        vec.swap(buf);
    };
    permute(s.grid_id, 1);
    permute(s.object_id, 1);
    permute(s.x0, 1);
    permute(s.x, 1);
    permute(s.v, 1);
    permute(s.V, 1);
    permute(s.m, 1);
    permute(s.rho, 1);
    permute(s.material, 1);
    permute(s.color, 3);
    permute(s.is_dynamic, 1);
    permute(s.uid, 1);
}

inline void add_wrench(SphHandle& s, int obj, V3 f, V3 t) {
    if (obj < 0 || obj >= SPH_MAX_OBJECTS) return;
#pragma omp critical(sph_wrench)
    {
        s.rigid_force[obj][0] += f.x; s.rigid_force[obj][1] += f.y; s.rigid_force[obj][2] += f.z;
        s.rigid_torque[obj][0] += t.x; s.rigid_torque[obj][1] += t.y; s.rigid_torque[obj][2] += t.z;
    }
}

// base_solver.py:105-123
void compute_rigid_particle_volume(SphHandle& s) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_RIGID) continue;
        if (!(s.x[i].y <= s.g_upper)) continue;
        float ret = s.kernel_W(0.0f);
        s.for_all_neighbors(i, [&](int j) {
            if (s.object_id[j] == s.object_id[i]) ret += s.kernel_W(norm(s.x[i] - s.x[j]));
        });
        s.V[i] = 1.0f / ret;
        s.m[i] = s.rho0 * s.V[i];
    }
}

// base_solver.py:521-541
void compute_density(SphHandle& s) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        float d = s.V[i] * s.kernel_W(0.0f);
        float ret = 0.f;
        s.for_all_neighbors(i, [&](int j) { ret += s.V[j] * s.kernel_W(norm(s.x[i] - s.x[j])); });
        d += ret;
        d *= s.rho0;
        s.rho[i] = d;
    }
}

// base_solver.py:135-187
void compute_pressure_acceleration(SphHandle& s) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.cap; i++) s.a[i] = {0, 0, 0};  // .fill(0.0) over the whole field
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (!s.is_dynamic[i]) continue;
        s.a[i] = {0, 0, 0};
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        V3 ret = {0, 0, 0};
        float den_i = s.rho[i];
        s.for_all_neighbors(i, [&](int j) {
            V3 R = s.x[i] - s.x[j];
            V3 nabla = s.kernel_gradient(R);
            if (s.material[j] == SPH_MATERIAL_FLUID) {
                float den_j = s.rho[j];
                ret += (-s.m[j] * (s.p[i] / (den_i * den_i) + s.p[j] / (den_j * den_j))) * nabla;
            } else if (s.material[j] == SPH_MATERIAL_RIGID) {
                V3 acc = (-s.rho0 * s.V[j] * s.p[i] / (den_i * den_i)) * nabla;
                ret += acc;
                if (s.is_dynamic[j]) {
                    int obj = s.object_id[j];
                    V3 force = ((s.rho0 * s.V[j] * s.p[i] / (den_i * den_i)) * nabla) * (s.rho0 * s.V[i]);
                    V3 com = (obj >= 0 && obj < SPH_MAX_OBJECTS) ? s.rigid_com[obj] : V3{0, 0, 0};
                    add_wrench(s, obj, force, cross(s.x[i] - com, force));  // arm uses x_i here (:185)
                }
            }
        });
        s.a[i] = ret;
    }
}

// base_solver.py:202-207
void compute_gravity_acceleration(SphHandle& s) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.a[i] = s.g;
}

// base_solver.py:209-229
void compute_surface_tension_acceleration(SphHandle& s) {
    const float diameter2 = (float)((double)s.P.dx * 2.0 * ((double)s.P.dx * 2.0));
    const float w_d = s.kernel_W(s.diameter);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        V3 a_i = {0, 0, 0};
        s.for_all_neighbors(i, [&](int j) {
            if (s.material[j] != SPH_MATERIAL_FLUID) return;
            V3 R = s.x[i] - s.x[j];
            float R2 = dot(R, R);
            float c = s.sigma / s.m[i] * s.m[j];
            if (R2 > diameter2)
                a_i -= (c * R) * s.kernel_W(norm(R));
            else
                a_i -= (c * R) * w_d;
        });
        s.a[i] += a_i;
    }
}

// base_solver.py:231-278
void compute_viscosity_acceleration_standard(SphHandle& s) {
    const float c_f = (float)(2.0 * (3 + 2) * s.P.viscosity);
    const float c_b = (float)(2.0 * (3 + 2) * s.P.viscosity_b);
    const float eps = (float)(0.01 * s.P.dh * s.P.dh);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        V3 a_i = {0, 0, 0};
        s.for_all_neighbors(i, [&](int j) {
            V3 R = s.x[i] - s.x[j];
            V3 nabla = s.kernel_gradient(R);
            float v_xy = dot(s.v[i] - s.v[j], R);
            float n = norm(R);
            if (s.material[j] == SPH_MATERIAL_FLUID) {
                float m_ij = (s.m[i] + s.m[j]) / 2;
                V3 acc = (c_f * m_ij / s.rho[j] / (n * n + eps) * v_xy) * nabla;
                a_i += acc;
            } else if (s.material[j] == SPH_MATERIAL_RIGID) {
                float m_ij = s.rho0 * s.V[j];
                V3 acc = (c_b * m_ij / s.rho[i] / (n * n + eps) * v_xy) * nabla;
                a_i += acc;
                if (s.is_dynamic[j]) {
                    int obj = s.object_id[j];
                    V3 force = (-acc * s.m[i]) / s.rho0;
                    V3 com = (obj >= 0 && obj < SPH_MAX_OBJECTS) ? s.rigid_com[obj] : V3{0, 0, 0};
                    add_wrench(s, obj, force, cross(s.x[j] - com, force));
                }
            }
        });
        s.a[i] += a_i / s.rho0;
    }
}

// ---- implicit viscosity (base_solver.py:281-517) ----
M3 compute_A_ij(const SphHandle& s, int i, int j) {  // :348-371
    const float eps = (float)(0.01 * s.P.dh * s.P.dh);
    M3 A = m3_zero();
    V3 R = s.x[i] - s.x[j];
    V3 nabla = s.kernel_gradient(R);
    if (s.material[j] == SPH_MATERIAL_FLUID) {
        float m_ij = (s.m[i] + s.m[j]) / 2;
        float c = (float)(-2.0 * (3 + 2) * s.P.viscosity) * m_ij / s.rho[j] / (norm_sqr(R) + eps);
        A = outer(nabla, R) * c;
    } else if (s.material[j] == SPH_MATERIAL_RIGID) {
        float m_ij = s.rho0 * s.V[j];
        float c = (float)(-2.0 * (3 + 2) * s.P.viscosity_b) * m_ij / s.rho[i] / (norm_sqr(R) + eps);
        A = outer(nabla, R) * c;
    }
    return A;
}

void cg_prepare1(SphHandle& s) {  // :281-315
    const float eps = (float)(0.01 * s.P.dh * s.P.dh);
    const float c_b = (float)(2.0 * (3 + 2) * s.P.viscosity_b);
    for (int i = 0; i < s.cap; i++) {
        s.cg_r[i] = s.cg_p[i] = s.v_orig[i] = s.cg_b[i] = s.cg_Ap[i] = {0, 0, 0};
    }
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.cg_x[i] += s.v[i];
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.v_orig[i] = s.v[i];
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        M3 ret = m3_zero();
        s.for_all_neighbors(i, [&](int j) { ret = ret - compute_A_ij(s, i, j); });  // compute_A_ii_task :325-331
        M3 diag = m3_identity() - (ret * s.dt) / s.rho0;
        s.cg_dinv[i] = inverse(diag);
        V3 ret1 = {0, 0, 0};
        s.for_all_neighbors(i, [&](int j) {  // compute_b_i_task :333-346
            if (s.material[j] != SPH_MATERIAL_RIGID) return;
            V3 R = s.x[i] - s.x[j];
            V3 nabla = s.kernel_gradient(R);
            ret1 += (c_b * s.rho0 * s.V[j] / s.rho[i] * dot(s.v[j], R) / (norm_sqr(R) + eps)) * nabla;
        });
        s.cg_b[i] = s.v[i] - (s.dt * ret1) / s.rho0;
        s.cg_p[i] = s.cg_x[i];
    }
}

void cg_prepare2(SphHandle& s) {  // :317-323
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            s.cg_r[i] = mv(s.cg_dinv[i], s.cg_b[i]) - s.cg_Ap[i];
            s.cg_p[i] = s.cg_r[i];
        }
}

void cg_compute_Ap(SphHandle& s) {  // :373-391
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        V3 ret = {0, 0, 0};
        s.for_all_neighbors(i, [&](int j) {
            if (s.material[j] != SPH_MATERIAL_FLUID) return;
            M3 A = compute_A_ij(s, i, j);
            ret += mv(mm(s.cg_dinv[i], -A), s.cg_p[j]);
        });
        ret = ret * s.dt;
        ret = ret / s.rho0;
        ret += s.cg_p[i];
        s.cg_Ap[i] = ret;
    }
}

void cg_compute_alpha(SphHandle& s) {  // :393-406
    double num = 0, den = 0;
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            num += norm_sqr(s.cg_r[i]);
            den += dot(s.cg_p[i], s.cg_Ap[i]);
        }
    float nf = (float)num, df = (float)den;
    s.cg_alpha = df > 1e-18f ? nf / df : 0.f;
}

void cg_update_x(SphHandle& s) {  // :408-412
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.cg_x[i] += s.cg_alpha * s.cg_p[i];
}

void cg_update_r_and_beta(SphHandle& s) {  // :414-431
    double num = 0, den = 0;
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            V3 nr = s.cg_r[i] - s.cg_alpha * s.cg_Ap[i];
            num += norm_sqr(nr);
            den += norm_sqr(s.cg_r[i]);
            s.cg_r[i] = nr;
        }
    float nf = (float)num, df = (float)den;
    s.cg_error = sqrtf(nf);
    s.cg_beta = df > 1e-18f ? nf / df : 0.f;
}

void cg_update_p(SphHandle& s) {  // :433-437
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.cg_p[i] = s.cg_r[i] + s.cg_beta * s.cg_p[i];
}

void cg_prepare_guess(SphHandle& s) {  // :439-443
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.cg_x[i] -= s.v_orig[i];
}

void viscosity_update_velocity(SphHandle& s) {  // :463-467
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.v[i] = s.cg_x[i];
}

void copy_back_original_velocity(SphHandle& s) {  // :469-473
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.v[i] = s.v_orig[i];
}

int conjugate_gradient_loop(SphHandle& s) {  // :445-461
    float tol = 1000.0f;
    int it = 0;
    while (tol > 1e-6f && it < 1000) {
        cg_compute_Ap(s);
        cg_compute_alpha(s);
        cg_update_x(s);
        cg_update_r_and_beta(s);
        cg_update_p(s);
        tol = s.cg_error;
        it++;
    }
    return it;
}

int implicit_viscosity_solve(SphHandle& s) {  // :509-517
    cg_prepare1(s);
    cg_compute_Ap(s);
    cg_prepare2(s);
    int it = conjugate_gradient_loop(s);
    viscosity_update_velocity(s);
    compute_viscosity_acceleration_standard(s);
    copy_back_original_velocity(s);
    cg_prepare_guess(s);
    return it;
}

// base_solver.py:190-200
int compute_non_pressure_acceleration(SphHandle& s, int* cg_it) {
    compute_gravity_acceleration(s);
    compute_surface_tension_acceleration(s);
    if (s.P.visc_method == SPH_VISC_STANDARD)
        compute_viscosity_acceleration_standard(s);
    else if (s.P.visc_method == SPH_VISC_IMPLICIT) {
        int it = implicit_viscosity_solve(s);
        if (cg_it) *cg_it = it;
    } else
        return SPH_E_UNSUPPORTED;
    return SPH_OK;
}

// base_solver.py:574-605 (+ simulate_collisions :544-549)
void enforce_domain_boundary_3D(SphHandle& s, int particle_type) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++) {
        if (!(s.material[i] == particle_type && s.is_dynamic[i])) continue;
        V3 pos = s.x[i];
        V3 n = {0, 0, 0};
        if (pos.x > s.dom.x - s.padding) { n.x += 1.0f; s.x[i].x = s.dom.x - s.padding; }
        if (pos.x <= s.padding) { n.x += -1.0f; s.x[i].x = s.padding; }
        if (pos.y > s.dom.y - s.padding) { n.y += 1.0f; s.x[i].y = s.dom.y - s.padding; }
        if (pos.y <= s.padding) { n.y += -1.0f; s.x[i].y = s.padding; }
        if (pos.z > s.dom.z - s.padding) { n.z += 1.0f; s.x[i].z = s.dom.z - s.padding; }
        if (pos.z <= s.padding) { n.z += -1.0f; s.x[i].z = s.padding; }
        float len = norm(n);
        if (len > 1e-6f) {
            V3 vec = n / len;
            const float c_f = 0.5f;
            s.v[i] -= ((1.0f + c_f) * dot(s.v[i], vec)) * vec;
        }
    }
}

// base_solver.py:615-629
void renew_rigid_particle_state(SphHandle& s) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++) {
        if (!(s.material[i] == SPH_MATERIAL_RIGID && s.is_dynamic[i])) continue;
        int obj = s.object_id[i];
        if (obj < 0 || obj >= SPH_MAX_OBJECTS) continue;
        if (!s.rigid_is_dynamic[obj]) continue;
        V3 q = s.x0[i] - s.rigid_com0[obj];
        V3 p = mv(s.rigid_rot[obj], q);
        s.x[i] = s.rigid_com[obj] + p;
        s.v[i] = s.rigid_vel[obj] + cross(s.rigid_omega[obj], p);
    }
}

// base_solver.py:642-649
void update_fluid_velocity(SphHandle& s) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.v[i] += s.dt * s.a[i];
}

// base_solver.py:651-666
void update_fluid_position(SphHandle& s) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            s.x[i] += s.dt * s.v[i];
        } else if (s.x[i].y > s.g_upper) {
            int obj = s.object_id[i];
            if (obj < 0 || obj >= SPH_MAX_OBJECTS) continue;  // App. B#3 guard
            if (s.object_material[obj] == SPH_MATERIAL_FLUID) {
                s.x[i] += s.dt * s.v[i];
                if (s.x[i].y <= s.g_upper) s.material[i] = SPH_MATERIAL_FLUID;
            }
        }
    }
}

// base_solver.py:669-677
void prepare_emitter(SphHandle& s) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID && s.x[i].y > s.g_upper) s.material[i] = SPH_MATERIAL_RIGID;
}

// ---- WCSPH.py:16-24 ----
void wcsph_compute_pressure(SphHandle& s) {
    const float gamma = 7.0f, stiffness = 50000.0f;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            float rho_i = std::max(s.rho[i], s.rho0);
            s.rho[i] = rho_i;
            s.p[i] = stiffness * (powf(rho_i / s.rho0, gamma) - 1.0f);
        }
}

// ---- DFSPH.py ----
void dfsph_compute_alpha(SphHandle& s) {  // :22-62
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        V3 grad_p_i = {0, 0, 0};
        float sum_grad_p_k = 0.f;
        s.for_all_neighbors(i, [&](int j) {
            V3 grad_p_j = (-s.V[j]) * s.kernel_gradient(s.x[i] - s.x[j]);
            if (s.material[j] == SPH_MATERIAL_FLUID) {
                sum_grad_p_k += norm_sqr(grad_p_j);
                grad_p_i += grad_p_j;
            } else if (s.material[j] == SPH_MATERIAL_RIGID) {
                grad_p_i += grad_p_j;
            }
        });
        sum_grad_p_k += norm_sqr(grad_p_i);
        s.alpha[i] = sum_grad_p_k > 1e-5f ? 1.0f / sum_grad_p_k : 0.0f;
    }
}

void dfsph_compute_density_derivative(SphHandle& s) {  // :65-101
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        float adv = 0.f;
        int nn = 0;
        s.for_all_neighbors(i, [&](int j) {
            adv += s.V[j] * dot(s.v[i] - s.v[j], s.kernel_gradient(s.x[i] - s.x[j]));
            nn += 1;
        });
        adv = std::max(adv, 0.0f);
        if (nn < 20) adv = 0.0f;
        s.drho[i] = adv;
    }
}

void dfsph_compute_density_star(SphHandle& s) {  // :104-126
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        float delta = 0.f;
        s.for_all_neighbors(i, [&](int j) { delta += s.V[j] * dot(s.v[i] - s.v[j], s.kernel_gradient(s.x[i] - s.x[j])); });
        float adv = s.rho[i] / s.rho0 + s.dt * delta;
        s.rho_star[i] = std::max(adv, 1.0f);
    }
}

void dfsph_compute_kappa_v(SphHandle& s) {  // :132-137
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.kappa_v[i] = s.drho[i] * s.alpha[i];
}

void dfsph_compute_kappa(SphHandle& s) {  // :217-223
    float dt_inv = 1 / s.dt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.kappa[i] = (s.rho_star[i] - 1.0f) * s.alpha[i] * dt_inv;
}

// correct_divergence_step :161-202 and correct_density_error_step :245-283 share one body; the
// in-place velocity update of the latter reads only kappa/rho/x/V of neighbours, so accumulating
// dv and adding once differs from upstream's repeated `v_i -= ...` only in rounding order; we
// follow each variant's own order.
void dfsph_correct_step(SphHandle& s, const std::vector<float>& kappa, bool accumulate_then_add) {
    const float m_eps = 1e-5f;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        float k_i = kappa[i];
        V3 dv = {0, 0, 0};
        V3 vi = s.v[i];
        s.for_all_neighbors(i, [&](int j) {
            V3 term = {0, 0, 0};
            bool hit = false;
            if (s.material[j] == SPH_MATERIAL_FLUID) {
                float k_j = kappa[j];
                float k_sum = k_i + k_j;
                if (fabsf(k_sum) > m_eps * s.dt) {
                    V3 grad_p_j = s.V[j] * s.kernel_gradient(s.x[i] - s.x[j]);
                    term = (grad_p_j * (k_i / s.rho[i] + k_j / s.rho[j])) * s.rho0;
                    hit = true;
                }
            } else if (s.material[j] == SPH_MATERIAL_RIGID) {
                float den_i = s.rho[i];
                if (fabsf(k_i) > m_eps * s.dt) {
                    V3 grad_p_j = s.V[j] * s.kernel_gradient(s.x[i] - s.x[j]);
                    term = (grad_p_j * (k_i / den_i)) * s.rho0;
                    hit = true;
                    if (s.is_dynamic[j]) {
                        int obj = s.object_id[j];
                        V3 force = ((grad_p_j * (k_i / den_i)) * s.rho0 / s.dt) * (s.V[i] * s.rho0);
                        V3 com = (obj >= 0 && obj < SPH_MAX_OBJECTS) ? s.rigid_com[obj] : V3{0, 0, 0};
                        add_wrench(s, obj, force, cross(s.x[j] - com, force));
                    }
                }
            }
            if (hit) {
                if (accumulate_then_add) dv -= term; else vi -= term;
            }
        });
        s.v[i] = accumulate_then_add ? s.v[i] + dv : vi;
    }
}

float dfsph_compute_density_derivative_error(SphHandle& s) {  // :205-211
    double e = 0;
#pragma omp parallel for schedule(static) reduction(+ : e)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) e += s.rho0 * s.drho[i];
    return (float)e / (float)s.N;
}

float dfsph_compute_density_error(SphHandle& s) {  // :285-294
    double e = 0;
#pragma omp parallel for schedule(static) reduction(+ : e)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) e += s.rho_star[i] - 1.0f;
    return (float)e / (float)s.N;
}

void dfsph_correct_divergence_error(SphHandle& s, int* iters, float* err) {  // :139-159
    const int max_it = 1000;
    const float max_error_V = 0.001f;
    int it = 0;
    dfsph_compute_density_derivative(s);
    float e = 0.f;
    while (it < 1 || it < max_it) {
        dfsph_compute_kappa_v(s);
        dfsph_correct_step(s, s.kappa_v, true);
        dfsph_compute_density_derivative(s);
        e = dfsph_compute_density_derivative_error(s);
        float eta = max_error_V * s.rho0 / s.dt;
        it++;
        if (e <= eta) break;
    }
    *iters = it;
    *err = e;
}

void dfsph_correct_density_error(SphHandle& s, int* iters, float* err) {  // :225-243
    const int max_it = 1000;
    const float max_error = 0.0001f;
    dfsph_compute_density_star(s);
    int it = 0;
    float e = 0.f;
    while (it < 1 || it < max_it) {
        dfsph_compute_kappa(s);
        dfsph_correct_step(s, s.kappa, false);
        dfsph_compute_density_star(s);
        e = dfsph_compute_density_error(s);
        it++;
        if (e <= max_error) break;
    }
    *iters = it;
    *err = e;
}

// ---- PCISPH.py ----
void pcisph_compute_predicted_velocity(SphHandle& s) {  // :18-22
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.v_pred[i] = s.v[i] + s.dt * (s.a[i] + s.a_p[i]);
}
void pcisph_compute_predicted_position(SphHandle& s) {  // :25-29
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) s.x_pred[i] = s.x[i] + s.dt * s.v_pred[i];
}
void pcisph_compute_density_star(SphHandle& s) {  // :32-62 (no self term; N(i) from current x)
    std::vector<float> part(s.N, 0.f);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        float ret = 0.f;
        V3 pos_i = s.x_pred[i];
        s.for_all_neighbors(i, [&](int j) {
            if (s.material[j] == SPH_MATERIAL_FLUID)
                ret += s.V[j] * s.kernel_W(norm(pos_i - s.x_pred[j]));
            else if (s.material[j] == SPH_MATERIAL_RIGID)
                ret += s.V[j] * s.kernel_W(norm(pos_i - s.x[j]));
        });
        s.rho_star[i] = ret * s.rho0;
        part[i] = std::max(0.0f, ret - 1.0f);
    }
    double e = 0;
    for (int i = 0; i < s.N; i++) e += part[i];
    s.density_error = s.Nfluid > 0 ? (float)e / (float)s.Nfluid : 0.f;
}
void pcisph_update_pressure(SphHandle& s) {  // :65-71
#pragma omp parallel for schedule(static)
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            s.p[i] += s.pcisph_k * (s.rho0 - s.rho_star[i]);
            if (s.p[i] < 0.0f) s.p[i] = 0.0f;
        }
}
void pcisph_compute_temp_pressure_acceleration(SphHandle& s) {  // :74-107
    for (int i = 0; i < s.cap; i++) s.a_p[i] = {0, 0, 0};
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < s.N; i++) {
        if (s.material[i] != SPH_MATERIAL_FLUID) continue;
        V3 ret = {0, 0, 0};
        float den_i = s.rho[i];
        s.for_all_neighbors(i, [&](int j) {
            V3 nabla = s.kernel_gradient(s.x[i] - s.x[j]);
            if (s.material[j] == SPH_MATERIAL_FLUID) {
                float den_j = s.rho[j];
                ret += (-s.m[j] * (s.p[i] / (den_i * den_i) + s.p[j] / (den_j * den_j))) * nabla;
            } else if (s.material[j] == SPH_MATERIAL_RIGID) {
                ret += (-s.rho0 * s.V[j] * (s.p[i] / (den_i * den_i))) * nabla;
            }
        });
        s.a_p[i] = ret;
    }
}
void pcisph_compute_k(SphHandle& s) {  // :128-151
    float support = s.h;
    float diam = (float)(s.P.dx * 2.0 * 0.97);
    V3 sumGradW = {0, 0, 0};
    float sumGradW2 = 0.f;
    int max_i = (int)(support / diam) + 1;
    for (int i = -max_i; i <= max_i; i++)
        for (int j = -max_i; j <= max_i; j++)
            for (int k = -max_i; k <= max_i; k++) {
                V3 pos_j = {i * diam, j * diam, k * diam};
                V3 x_ij = V3{0, 0, 0} - pos_j;
                if (norm(x_ij) < support) {
                    V3 nabla = s.kernel_gradient(x_ij);
                    sumGradW += nabla;
                    sumGradW2 += norm_sqr(nabla);
                }
            }
    s.pcisph_k = -0.5f / (s.dt * s.V0) / (s.dt * s.V0) / (norm_sqr(sumGradW) + sumGradW2);
}
void pcisph_init_step(SphHandle& s) {  // :153-162
    for (int i = 0; i < s.cap; i++) { s.a_p[i] = {0, 0, 0}; s.p[i] = 0.f; }
    s.density_error = 100.0f;
    for (int i = 0; i < s.N; i++)
        if (s.material[i] == SPH_MATERIAL_FLUID) {
            s.v_pred[i] = s.v[i] + s.dt * s.a[i];
            s.x_pred[i] = s.x[i] + s.dt * s.v_pred[i];
        }
}
void pcisph_refine(SphHandle& s, int* iters, float* err) {  // :110-125
    int it = 0;
    while (it < 1000) {
        pcisph_compute_density_star(s);
        pcisph_update_pressure(s);
        pcisph_compute_temp_pressure_acceleration(s);
        pcisph_compute_predicted_velocity(s);
        pcisph_compute_predicted_position(s);
        it++;
        if (s.density_error < 0.001f) break;
    }
    *iters = it;
    *err = s.density_error;
}

// solver _step bodies with the rigid-solver / insert_object hooks being no-ops
int step_once(SphHandle& s, SphStepStats* st) {
    int cg_it = 0, rc;
    if (s.P.method == SPH_METHOD_WCSPH) {  // WCSPH.py:27-45
        prepare_neighborhood_search(s);
        compute_density(s);
        if ((rc = compute_non_pressure_acceleration(s, &cg_it))) return rc;
        update_fluid_velocity(s);
        wcsph_compute_pressure(s);
        compute_pressure_acceleration(s);
        update_fluid_velocity(s);
        update_fluid_position(s);
        renew_rigid_particle_state(s);
        enforce_domain_boundary_3D(s, SPH_MATERIAL_FLUID);
    } else if (s.P.method == SPH_METHOD_PCISPH) {  // PCISPH.py:165-185
        prepare_neighborhood_search(s);
        compute_density(s);
        if ((rc = compute_non_pressure_acceleration(s, &cg_it))) return rc;
        pcisph_init_step(s);
        int it; float e;
        pcisph_refine(s, &it, &e);
        st->pcisph_iterations = it; st->pcisph_density_error = e; st->total_pcisph_iterations += it;
        update_fluid_velocity(s);
        compute_pressure_acceleration(s);
        update_fluid_velocity(s);
        update_fluid_position(s);
        renew_rigid_particle_state(s);
        enforce_domain_boundary_3D(s, SPH_MATERIAL_FLUID);
    } else if (s.P.method == SPH_METHOD_DFSPH) {  // DFSPH.py:298-319
        if ((rc = compute_non_pressure_acceleration(s, &cg_it))) return rc;
        update_fluid_velocity(s);
        int it; float e;
        dfsph_correct_density_error(s, &it, &e);
        st->dfsph_iterations = it; st->dfsph_density_error = e; st->total_dfsph_iterations += it;
        update_fluid_position(s);
        renew_rigid_particle_state(s);
        enforce_domain_boundary_3D(s, SPH_MATERIAL_FLUID);
        prepare_neighborhood_search(s);
        compute_density(s);
        dfsph_compute_alpha(s);
        dfsph_correct_divergence_error(s, &it, &e);
        st->dfsph_iterations_v = it; st->dfsph_divergence_error = e; st->total_dfsph_iterations_v += it;
    } else
        return SPH_E_UNSUPPORTED;
    st->cg_iterations = cg_it; st->cg_error = s.cg_error; st->total_cg_iterations += cg_it;
    // BaseSolver.step tail (base_solver.py:692-696)
    compute_rigid_particle_volume(s);
    return SPH_OK;
}

struct FieldRef { void* ptr; int comps; bool is_float; };

FieldRef field_ref(SphHandle& s, int f) {
    switch (f) {
        case SPH_F_OBJECT_ID: return {s.object_id.data(), 1, false};
        case SPH_F_POSITION: return {s.x.data(), 3, true};
        case SPH_F_VELOCITY: return {s.v.data(), 3, true};
        case SPH_F_ACCELERATION: return {s.a.data(), 3, true};
        case SPH_F_REST_VOLUME: return {s.V.data(), 1, true};
        case SPH_F_MASS: return {s.m.data(), 1, true};
        case SPH_F_DENSITY: return {s.rho.data(), 1, true};
        case SPH_F_PRESSURE: return {s.p.data(), 1, true};
        case SPH_F_MATERIAL: return {s.material.data(), 1, false};
        case SPH_F_COLOR: return {s.color.data(), 3, false};
        case SPH_F_IS_DYNAMIC: return {s.is_dynamic.data(), 1, false};
        case SPH_F_ORIGINAL_POSITION: return {s.x0.data(), 3, true};
        case SPH_F_GRID_ID: return {s.grid_id.data(), 1, false};
        case SPH_F_UID: return {s.uid.data(), 1, false};
        case SPH_F_DFSPH_ALPHA: return {s.alpha.data(), 1, true};
        case SPH_F_DFSPH_KAPPA: return {s.kappa.data(), 1, true};
        case SPH_F_DFSPH_KAPPA_V: return {s.kappa_v.data(), 1, true};
        case SPH_F_DENSITY_STAR: return {s.rho_star.data(), 1, true};
        case SPH_F_DENSITY_DERIVATIVE: return {s.drho.data(), 1, true};
        case SPH_F_PRESSURE_ACCELERATION: return {s.a_p.data(), 3, true};
        case SPH_F_PREDICTED_VELOCITY: return {s.v_pred.data(), 3, true};
        case SPH_F_PREDICTED_POSITION: return {s.x_pred.data(), 3, true};
        case SPH_F_CG_P: return {s.cg_p.data(), 3, true};
        case SPH_F_ORIGINAL_VELOCITY: return {s.v_orig.data(), 3, true};
        case SPH_F_CG_AP: return {s.cg_Ap.data(), 3, true};
        case SPH_F_CG_X: return {s.cg_x.data(), 3, true};
        case SPH_F_CG_B: return {s.cg_b.data(), 3, true};
        case SPH_F_CG_R: return {s.cg_r.data(), 3, true};
        case SPH_F_CG_DIAG_INV: return {s.cg_dinv.data(), 9, true};
        default: return {nullptr, 0, false};
    }
}

}  // namespace

extern "C" {

int sph_abi_version(void) { return SPH_ABI_VERSION; }
const char* sph_backend_name(void) { return "oracle-cpu"; }
const char* sph_last_error(const SphHandle* h) { return h ? h->err.c_str() : "null handle"; }

// CPU-baseline fairness on multi-socket hosts: every array is first touched by the creating thread, which would
// put all particle data on one NUMA node while the OpenMP team spans all of them.  Interleave this process's
// future pages over the online nodes instead (MPOL_INTERLEAVE; silently skipped where the syscall is filtered,
// on single-node hosts, or with SPH_ORACLE_NO_INTERLEAVE=1).
static void interleave_memory_once() {
    static bool done = false;
    if (done) return;
    done = true;
    if (std::getenv("SPH_ORACLE_NO_INTERLEAVE")) return;
    FILE* f = std::fopen("/sys/devices/system/node/online", "r");
    if (!f) return;
    int lo = 0, hi = 0;
    int n = std::fscanf(f, "%d-%d", &lo, &hi);
    std::fclose(f);
    if (n < 2 || hi <= lo || hi >= 1024) return;            // "0" (one node) or an unexpected format
    unsigned long mask[16] = {0};
    for (int node = lo; node <= hi; node++) mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
#ifdef SYS_set_mempolicy
    (void)syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, mask, (unsigned long)(hi + 2));
#endif
}

int sph_create(const SphParams* p, SphHandle** out) {
    if (!p || !out) return SPH_E_INVALID;
    interleave_memory_once();
    if (p->abi_version != SPH_ABI_VERSION || p->dim != 3 || p->max_particles < 0) return SPH_E_INVALID;
    SphHandle* s = new SphHandle();
    s->P = *p;
    s->cap = p->max_particles;
    s->ncell = p->grid_num[0] * p->grid_num[1] * p->grid_num[2];
    s->h = (float)p->dh; s->dx = (float)p->dx; s->diameter = (float)(2.0 * p->dx); s->V0 = (float)p->V0;
    s->rho0 = (float)p->density0; s->dt = (float)p->dt; s->g_upper = (float)p->g_upper;
    s->visc = (float)p->viscosity; s->visc_b = (float)p->viscosity_b; s->sigma = (float)p->surface_tension;
    s->padding = (float)p->padding;
    s->g = {(float)p->gravity[0], (float)p->gravity[1], (float)p->gravity[2]};
    s->dom = {(float)p->domain_size[0], (float)p->domain_size[1], (float)p->domain_size[2]};
    double k = 8.0 / M_PI / (p->dh * p->dh * p->dh);
    s->kW = (float)k; s->kW2 = (float)(k * 2); s->kG = (float)(6.0 * k);
    size_t n = (size_t)s->cap;
    s->object_id.assign(n, 0); s->material.assign(n, 0); s->is_dynamic.assign(n, 0); s->grid_id.assign(n, 0);
    s->uid.assign(n, 0); s->color.assign(n * 3, 0);
    V3 z = {0, 0, 0};
    s->x.assign(n, z); s->v.assign(n, z); s->a.assign(n, z); s->x0.assign(n, z);
    s->V.assign(n, 0.f); s->m.assign(n, 0.f); s->rho.assign(n, 0.f); s->p.assign(n, 0.f);
    s->alpha.assign(n, 0.f); s->kappa.assign(n, 0.f); s->kappa_v.assign(n, 0.f); s->rho_star.assign(n, 0.f); s->drho.assign(n, 0.f);
    s->a_p.assign(n, z); s->v_pred.assign(n, z); s->x_pred.assign(n, z);
    s->cg_p.assign(n, z); s->v_orig.assign(n, z); s->cg_Ap.assign(n, z); s->cg_x.assign(n, z); s->cg_b.assign(n, z); s->cg_r.assign(n, z);
    s->cg_dinv.assign(n, m3_zero());
    s->cell_scan.assign((size_t)s->ncell, 0);
    for (int o = 0; o < SPH_MAX_OBJECTS; o++) {
        s->rigid_com0[o] = s->rigid_com[o] = s->rigid_vel[o] = s->rigid_omega[o] = z;
        s->rigid_rot[o] = m3_zero();
        for (int d = 0; d < 3; d++) s->rigid_force[o][d] = s->rigid_torque[o][d] = 0;
    }
    *out = s;
    return SPH_OK;
}

int sph_destroy(SphHandle* h) { delete h; return SPH_OK; }

int sph_add_particles(SphHandle* h, int32_t object_id, int32_t n, const float* x, const float* v, const float* density,
                      const float* pressure, const int32_t* material, const int32_t* is_dynamic, const int32_t* color) {
    if (!h || n < 0) return SPH_E_INVALID;
    if (h->N + n > h->cap) return fail(h, SPH_E_CAPACITY, "particle_max_num exceeded");
    for (int k = 0; k < n; k++) {  // add_particle, base_container.py:403-415
        int p = h->N + k;
        h->object_id[p] = object_id;
        h->x[p] = {x[3 * k], x[3 * k + 1], x[3 * k + 2]};
        h->x0[p] = h->x[p];
        h->v[p] = {v[3 * k], v[3 * k + 1], v[3 * k + 2]};
        h->rho[p] = density[k];
        h->V[p] = h->V0;
        h->m[p] = h->V0 * density[k];
        h->p[p] = pressure[k];
        h->material[p] = material[k];
        h->is_dynamic[p] = is_dynamic[k];
        for (int c = 0; c < 3; c++) h->color[3 * p + c] = color[3 * k + c];
        h->uid[p] = p;
    }
    h->N += n;
    return SPH_OK;
}

int sph_get_field(SphHandle* h, int32_t field, void* dst, size_t bytes) {
    if (!h || !dst) return SPH_E_INVALID;
    if (field == SPH_F_CELL) {
        if (bytes > (size_t)h->N * 12) return fail(h, SPH_E_INVALID, "size");
        int32_t* d = (int32_t*)dst;
        for (size_t i = 0; i < bytes / 12; i++) { int c[3]; h->pos_to_index(h->x[i], c); d[3 * i] = c[0]; d[3 * i + 1] = c[1]; d[3 * i + 2] = c[2]; }
        return SPH_OK;
    }
    if (field == SPH_F_NEIGHBOR_COUNT) {
        if (bytes > (size_t)h->N * 4) return fail(h, SPH_E_INVALID, "size");
        int32_t* d = (int32_t*)dst;
        for (size_t i = 0; i < bytes / 4; i++) { int c = 0; h->for_all_neighbors((int)i, [&](int) { c++; }); d[i] = c; }
        return SPH_OK;
    }
    FieldRef r = field_ref(*h, field);
    if (!r.ptr) return fail(h, SPH_E_INVALID, "unknown field");
    if (bytes > (size_t)h->cap * r.comps * 4) return fail(h, SPH_E_INVALID, "size exceeds field");
    memcpy(dst, r.ptr, bytes);
    return SPH_OK;
}

int sph_set_field(SphHandle* h, int32_t field, const void* src, size_t bytes) {
    if (!h || !src) return SPH_E_INVALID;
    FieldRef r = field_ref(*h, field);
    if (!r.ptr) return fail(h, SPH_E_INVALID, "unknown field");
    if (bytes > (size_t)h->cap * r.comps * 4) return fail(h, SPH_E_INVALID, "size exceeds field");
    memcpy(r.ptr, src, bytes);
    return SPH_OK;
}

int sph_fill_field(SphHandle* h, int32_t field, double value) {
    if (!h) return SPH_E_INVALID;
    FieldRef r = field_ref(*h, field);
    if (!r.ptr) return fail(h, SPH_E_INVALID, "unknown field");
    size_t n = (size_t)h->cap * r.comps;
    if (r.is_float) { float* p = (float*)r.ptr; for (size_t i = 0; i < n; i++) p[i] = (float)value; }
    else { int32_t* p = (int32_t*)r.ptr; for (size_t i = 0; i < n; i++) p[i] = (int32_t)value; }
    return SPH_OK;
}

int sph_field_ptr(SphHandle* h, int32_t field, void** ptr, int32_t* stride, int32_t* comps) {
    if (!h) return SPH_E_INVALID;
    FieldRef r = field_ref(*h, field);
    if (!r.ptr) return fail(h, SPH_E_INVALID, "unknown field");
    if (ptr) *ptr = r.ptr;
    if (stride) *stride = r.comps * 4;
    if (comps) *comps = r.comps;
    return SPH_OK;
}

int sph_get_scalar(SphHandle* h, int32_t s, double* out) {
    if (!h || !out) return SPH_E_INVALID;
    switch (s) {
        case SPH_S_DT: *out = h->dt; break;
        case SPH_S_PARTICLE_NUM: *out = h->N; break;
        case SPH_S_FLUID_PARTICLE_NUM: *out = h->Nfluid; break;
        case SPH_S_PCISPH_K: *out = h->pcisph_k; break;
        case SPH_S_DENSITY_ERROR: *out = h->density_error; break;
        case SPH_S_CG_ALPHA: *out = h->cg_alpha; break;
        case SPH_S_CG_BETA: *out = h->cg_beta; break;
        case SPH_S_CG_ERROR: *out = h->cg_error; break;
        case SPH_S_G_UPPER: *out = h->g_upper; break;
        case SPH_S_VISCOSITY: *out = h->P.viscosity; break;
        case SPH_S_VISCOSITY_B: *out = h->P.viscosity_b; break;
        case SPH_S_NUM_CELLS: *out = h->ncell; break;
        case SPH_S_MAX_PARTICLES: *out = h->cap; break;
        case SPH_S_ACTIVE_BRICKS: case SPH_S_MAX_WINDOW_SLOTS: case SPH_S_WINDOW_OVERFLOWS: *out = 0; break;   // CUDA-path diagnostics
        default: return fail(h, SPH_E_INVALID, "unknown scalar");
    }
    return SPH_OK;
}

int sph_set_scalar(SphHandle* h, int32_t s, double v) {
    if (!h) return SPH_E_INVALID;
    switch (s) {
        case SPH_S_DT: h->dt = (float)v; h->P.dt = v; break;
        case SPH_S_PARTICLE_NUM: if (v < 0 || v > h->cap) return fail(h, SPH_E_CAPACITY, "particle_num"); h->N = (int)v; break;
        case SPH_S_FLUID_PARTICLE_NUM: h->Nfluid = (int)v; break;
        case SPH_S_PCISPH_K: h->pcisph_k = (float)v; break;
        case SPH_S_DENSITY_ERROR: h->density_error = (float)v; break;
        case SPH_S_CG_ALPHA: h->cg_alpha = (float)v; break;
        case SPH_S_CG_BETA: h->cg_beta = (float)v; break;
        case SPH_S_CG_ERROR: h->cg_error = (float)v; break;
        case SPH_S_G_UPPER: h->g_upper = (float)v; h->P.g_upper = v; break;
        case SPH_S_VISCOSITY: h->P.viscosity = v; h->visc = (float)v; break;
        case SPH_S_VISCOSITY_B: h->P.viscosity_b = v; h->visc_b = (float)v; break;
        default: return fail(h, SPH_E_INVALID, "scalar not settable");
    }
    return SPH_OK;
}

int sph_set_object(SphHandle* h, int32_t obj, int32_t material, int32_t is_dynamic) {
    if (!h || obj < 0 || obj >= SPH_MAX_OBJECTS) return SPH_E_INVALID;
    h->object_material[obj] = material;
    h->rigid_is_dynamic[obj] = is_dynamic;
    return SPH_OK;
}

int sph_set_rigid_state(SphHandle* h, int32_t obj, const float com0[3], const float com[3], const float rot[9],
                        const float vel[3], const float omega[3]) {
    if (!h || obj < 0 || obj >= SPH_MAX_OBJECTS) return SPH_E_INVALID;
    if (com0) h->rigid_com0[obj] = {com0[0], com0[1], com0[2]};
    if (com) h->rigid_com[obj] = {com[0], com[1], com[2]};
    if (rot) memcpy(h->rigid_rot[obj].m, rot, 36);
    if (vel) h->rigid_vel[obj] = {vel[0], vel[1], vel[2]};
    if (omega) h->rigid_omega[obj] = {omega[0], omega[1], omega[2]};
    return SPH_OK;
}

int sph_get_rigid_wrench(SphHandle* h, float* force, float* torque) {
    if (!h) return SPH_E_INVALID;
    for (int o = 0; o < SPH_MAX_OBJECTS; o++)
        for (int d = 0; d < 3; d++) {
            if (force) force[3 * o + d] = (float)h->rigid_force[o][d];
            if (torque) torque[3 * o + d] = (float)h->rigid_torque[o][d];
        }
    return SPH_OK;
}

int sph_zero_rigid_wrench(SphHandle* h) {
    if (!h) return SPH_E_INVALID;
    for (int o = 0; o < SPH_MAX_OBJECTS; o++)
        for (int d = 0; d < 3; d++) h->rigid_force[o][d] = h->rigid_torque[o][d] = 0;
    return SPH_OK;
}

int sph_compute_rigid_body_mass(SphHandle* h, int32_t object_id, float* out) {  // base_container.py:384-390
    if (!h || !out) return SPH_E_INVALID;
    double sum = 0;
    for (int i = 0; i < h->N; i++)
        if (h->object_id[i] == object_id && h->is_dynamic[i]) sum += h->rho[i] * h->V0;
    *out = (float)sum;
    return SPH_OK;
}

int sph_prepare_neighborhood_search(SphHandle* h) {
    if (!h) return SPH_E_INVALID;
    prepare_neighborhood_search(*h);
    return SPH_OK;
}

int sph_get_neighbors(SphHandle* h, int32_t* offsets, int32_t* indices, size_t capacity) {
    if (!h || !offsets) return SPH_E_INVALID;
    size_t total = 0;
    for (int i = 0; i < h->N; i++) {
        offsets[i] = (int32_t)total;
        h->for_all_neighbors(i, [&](int j) {
            if (indices && total < capacity) indices[total] = j;
            total++;
        });
    }
    offsets[h->N] = (int32_t)total;
    if (indices && total > capacity) return fail(h, SPH_E_CAPACITY, "indices capacity too small");
    return SPH_OK;
}

int sph_get_grid_num_particles(SphHandle* h, int32_t* dst, size_t count) {
    if (!h || !dst || count > (size_t)h->ncell) return SPH_E_INVALID;
    memcpy(dst, h->cell_scan.data(), count * 4);
    return SPH_OK;
}

int sph_run_task(SphHandle* h, int32_t task, int32_t iarg, float* out) {
    if (!h) return SPH_E_INVALID;
    SphHandle& s = *h;
    switch (task) {
        case SPH_T_COMPUTE_RIGID_PARTICLE_VOLUME: compute_rigid_particle_volume(s); break;
        case SPH_T_COMPUTE_PRESSURE_ACCELERATION: compute_pressure_acceleration(s); break;
        case SPH_T_COMPUTE_GRAVITY_ACCELERATION: compute_gravity_acceleration(s); break;
        case SPH_T_COMPUTE_SURFACE_TENSION_ACCELERATION: compute_surface_tension_acceleration(s); break;
        case SPH_T_COMPUTE_VISCOSITY_ACCELERATION_STANDARD: compute_viscosity_acceleration_standard(s); break;
        case SPH_T_COMPUTE_DENSITY: compute_density(s); break;
        case SPH_T_ENFORCE_DOMAIN_BOUNDARY_3D: enforce_domain_boundary_3D(s, iarg); break;
        case SPH_T_RENEW_RIGID_PARTICLE_STATE: renew_rigid_particle_state(s); break;
        case SPH_T_UPDATE_FLUID_VELOCITY: update_fluid_velocity(s); break;
        case SPH_T_UPDATE_FLUID_POSITION: update_fluid_position(s); break;
        case SPH_T_PREPARE_EMITTER: prepare_emitter(s); break;
        case SPH_T_INIT_OBJECT_ID: std::fill(s.object_id.begin(), s.object_id.end(), -1); break;
        case SPH_T_INIT_ACCELERATION: std::fill(s.a.begin(), s.a.end(), V3{0, 0, 0}); break;
        case SPH_T_INIT_RIGID_BODY_FORCE_AND_TORQUE: sph_zero_rigid_wrench(h); break;
        case SPH_T_CG_PREPARE1: cg_prepare1(s); break;
        case SPH_T_CG_PREPARE2: cg_prepare2(s); break;
        case SPH_T_CG_COMPUTE_AP: cg_compute_Ap(s); break;
        case SPH_T_CG_COMPUTE_ALPHA: cg_compute_alpha(s); break;
        case SPH_T_CG_UPDATE_X: cg_update_x(s); break;
        case SPH_T_CG_UPDATE_R_AND_BETA: cg_update_r_and_beta(s); if (out) *out = s.cg_error; break;
        case SPH_T_CG_UPDATE_P: cg_update_p(s); break;
        case SPH_T_CG_PREPARE_GUESS: cg_prepare_guess(s); break;
        case SPH_T_VISCOSITY_UPDATE_VELOCITY: viscosity_update_velocity(s); break;
        case SPH_T_COPY_BACK_ORIGINAL_VELOCITY: copy_back_original_velocity(s); break;
        case SPH_T_WCSPH_COMPUTE_PRESSURE: wcsph_compute_pressure(s); break;
        case SPH_T_DFSPH_COMPUTE_ALPHA: dfsph_compute_alpha(s); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE: dfsph_compute_density_derivative(s); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_STAR: dfsph_compute_density_star(s); break;
        case SPH_T_DFSPH_COMPUTE_KAPPA_V: dfsph_compute_kappa_v(s); break;
        case SPH_T_DFSPH_CORRECT_DIVERGENCE_STEP: dfsph_correct_step(s, s.kappa_v, true); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_DERIVATIVE_ERROR: { float e = dfsph_compute_density_derivative_error(s); if (out) *out = e; break; }
        case SPH_T_DFSPH_COMPUTE_KAPPA: dfsph_compute_kappa(s); break;
        case SPH_T_DFSPH_CORRECT_DENSITY_ERROR_STEP: dfsph_correct_step(s, s.kappa, false); break;
        case SPH_T_DFSPH_COMPUTE_DENSITY_ERROR: { float e = dfsph_compute_density_error(s); if (out) *out = e; break; }
        case SPH_T_PCISPH_COMPUTE_PREDICTED_VELOCITY: pcisph_compute_predicted_velocity(s); break;
        case SPH_T_PCISPH_COMPUTE_PREDICTED_POSITION: pcisph_compute_predicted_position(s); break;
        case SPH_T_PCISPH_COMPUTE_DENSITY_STAR: pcisph_compute_density_star(s); if (out) *out = s.density_error; break;
        case SPH_T_PCISPH_UPDATE_PRESSURE: pcisph_update_pressure(s); break;
        case SPH_T_PCISPH_COMPUTE_TEMP_PRESSURE_ACCELERATION: pcisph_compute_temp_pressure_acceleration(s); break;
        case SPH_T_PCISPH_COMPUTE_K: pcisph_compute_k(s); if (out) *out = s.pcisph_k; break;
        case SPH_T_PCISPH_INIT_STEP: pcisph_init_step(s); break;
        default: return fail(h, SPH_E_INVALID, "unknown task");
    }
    return SPH_OK;
}

int sph_step(SphHandle* h, int32_t n_steps, SphStepStats* stats) {
    if (!h || n_steps < 0) return SPH_E_INVALID;
    SphStepStats st;
    memset(&st, 0, sizeof st);
    for (int k = 0; k < n_steps; k++) {
        int rc = step_once(*h, &st);
        if (rc) return fail(h, rc, "step failed");
        st.steps++;
    }
    if (stats) *stats = st;
    return SPH_OK;
}

int sph_dfsph_correct_density_error(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    int i; float err;
    dfsph_correct_density_error(*h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return SPH_OK;
}
int sph_dfsph_correct_divergence_error(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    int i; float err;
    dfsph_correct_divergence_error(*h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return SPH_OK;
}
int sph_pcisph_refine(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    int i; float err;
    pcisph_refine(*h, &i, &err);
    if (it) *it = i;
    if (e) *e = err;
    return SPH_OK;
}
int sph_implicit_viscosity_solve(SphHandle* h, int32_t* it, float* e) {
    if (!h) return SPH_E_INVALID;
    int i = implicit_viscosity_solve(*h);
    if (it) *it = i;
    if (e) *e = h->cg_error;
    return SPH_OK;
}

int sph_synchronize(SphHandle*) { return SPH_OK; }
int sph_set_stream(SphHandle*, void*) { return SPH_OK; }
int sph_profile_enable(SphHandle*, int32_t) { return SPH_OK; }
int sph_profile_read(SphHandle*, SphKernelStat*, int32_t, int32_t* count) { if (count) *count = 0; return SPH_OK; }

// Z-slab sharding is a property of the CUDA product; the oracle always holds the whole domain.
int sph_slab_unique_id(void*) { return SPH_E_UNSUPPORTED; }
int sph_slab_init(SphHandle* h, int32_t, int32_t, const void*, int32_t, int32_t, int64_t) { return fail(h, SPH_E_UNSUPPORTED, "oracle holds the whole domain"); }
int sph_slab_set_global_particle_num(SphHandle* h, int64_t) { return fail(h, SPH_E_UNSUPPORTED, "oracle"); }
int sph_slab_info(SphHandle* h, SphSlabInfo*) { return fail(h, SPH_E_UNSUPPORTED, "oracle"); }
int sph_slab_peer_export(SphHandle* h, void*) { return fail(h, SPH_E_UNSUPPORTED, "oracle"); }
int sph_slab_peer_import(SphHandle* h, int32_t, const void*) { return fail(h, SPH_E_UNSUPPORTED, "oracle"); }

}  // extern "C"
