// sph_window.cuh — shared-memory neighbour windows staged by TMA bulk copies.
//
// A CTA owns a chunk of SPH_BLOCK consecutive particles of the cell-sorted SoA.  Because the
// flatten is x-fastest, the 27-cell windows of all particles of the chunk are covered by at most 9
// contiguous index ranges of the sorted arrays (one per (dz, dy) row offset, overlapping ranges
// merged).  One elected thread issues cp.async.bulk (TMA, SASS UBLKCP) copies of those ranges of
// every payload array the task reads (all payloads are float4 per particle, so any range is
// 16-byte aligned) and the CTA waits on one mbarrier.  Afterwards every neighbour access is an
// LDS.128 at a 16-bit window index: the neighbour lists store those indices (2 B per pair), and the
// gathers that limited the global-memory version (one L1 wavefront per distinct 128-byte line per
// load) become shared-memory reads.
//
// Descriptor per chunk (written once per sort by k_chunk_windows, 40 ints):
//   [0] total window entries   [1] number of merged copies   [2..3] reserved
//   [4..12]  gstart[r]  first sorted index of row-offset r's range
//   [13..21] sbase[r]   window index of that first particle
//   [22..30] cp_g[c]    merged copy c: first sorted index
//   [31..39] cp_n[c]    merged copy c: particle count
#pragma once

#include "sph_common.cuh"

#define SPH_DESC_INTS 40

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// payload views: the same task body reads neighbour data either from the staged window (index =
// window slot) or straight from global memory (index = sorted particle index; unstaged chunks)
template <int NPAY>
struct SmemView {
    const float4* a[NPAY];
    __device__ __forceinline__ float4 get(int which, int idx) const { return a[which][idx]; }
};
template <int NPAY>
struct GlobalView {
    const float4* a[NPAY];
    __device__ __forceinline__ float4 get(int which, int idx) const { return __ldg(a[which] + idx); }
};

template <int NPAY>
struct Window {
    SmemView<NPAY> sv;
    GlobalView<NPAY> gv;
    const int* desc;     // shared copy of the chunk descriptor
    bool staged;
    // sorted particle index of a window slot (rare paths: rigid wrench, debug)
    __device__ __forceinline__ int global_index(int w) const {
        const int ncp = desc[1];
        int s = 0;
        for (int c = 0; c < ncp; c++) {
            const int n = desc[31 + c];
            if (w < s + n) return desc[22 + c] + (w - s);
            s += n;
        }
        return -1;
    }
};

// Stage the chunk's window: call from all threads of the CTA, before any divergence.
// payload[k] are the global float4 arrays; smem must hold wmax * NPAY float4.
template <int NPAY>
__device__ __forceinline__ Window<NPAY> window_open(const Dev& d, const float4* const (&payload)[NPAY], float4* smem, int wmax,
                                                    int* s_desc, unsigned long long* s_mbar) {
    const int tid = threadIdx.x;
    if (tid < SPH_DESC_INTS) s_desc[tid] = d.chunk_desc[(size_t)blockIdx.x * SPH_DESC_INTS + tid];
    if (tid == 0) {
        mbar_init(s_mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    Window<NPAY> w;
    w.desc = s_desc;
    const int total = s_desc[0];
    // chunks without fluid rows (as of the last sort) have nothing to sum: do not stage; a row that
    // turned fluid since (emitter) walks global memory instead
    w.staged = s_desc[2] > 0 && total <= wmax;
#pragma unroll
    for (int k = 0; k < NPAY; k++) {
        w.sv.a[k] = smem + (size_t)k * wmax;
        w.gv.a[k] = payload[k];
    }
    if (w.staged) {
        if (tid == 0) {
            mbar_arrive_expect_tx(s_mbar, (unsigned)total * 16u * NPAY);
            const int ncp = s_desc[1];
            int s = 0;
            for (int c = 0; c < ncp; c++) {
                const int g = s_desc[22 + c], n = s_desc[31 + c];
#pragma unroll
                for (int k = 0; k < NPAY; k++)
                    tma_bulk_load(smem + (size_t)k * wmax + s, payload[k] + g, (unsigned)n * 16u, s_mbar);
                s += n;
            }
        }
        mbar_wait(s_mbar, 0);
    }
    return w;
}

// Walk the 27-cell window of particle i with candidates read from the staged window.
// visit(view, idx, pj, R, r2): idx is a window slot (view = smem) — accepted neighbours only.
template <int NPAY, class Visit>
__device__ __forceinline__ void window_walk(const Consts& c, const Dev& d, const Window<NPAY>& win, int i, float4 pi, Visit&& visit) {
    const int3 g = cell_of(c, pi.x, pi.y, pi.z);
    const int xlo = max(g.x - 1, 0), xhi = min(g.x + 1, c.nx - 1);
    const float4* __restrict__ s_pv = win.sv.a[0];
#pragma unroll 1
    for (int dz = -1; dz <= 1; dz++) {
        const int zz = g.z + dz;
        if (zz < 0 || zz >= c.nz) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            const int yy = g.y + dy;
            if (yy < 0 || yy >= c.ny) continue;
            const int r = (dz + 1) * 3 + (dy + 1);
            const int row = (zz * c.ny + yy) * c.nx;
            const int s = __ldg(d.cell_start + row + xlo);
            const int e = __ldg(d.cell_start + row + xhi + 1);
            int w = win.desc[13 + r] + (s - win.desc[4 + r]);
#pragma unroll 1
            for (int j = s; j < e; j++, w++) {
                const float4 pj = s_pv[w];
                const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                const float r2 = dist2(R);
                if (r2 < c.h2_thresh && j != i) visit(win.sv, w, pj, R, r2);
            }
        }
    }
}

// Stream the 16-bit neighbour list of particle i; positions come from the staged window.
template <int NPAY, class Visit>
__device__ __forceinline__ void window_list(const Dev& d, const Window<NPAY>& win, int i, int n, float4 pi, Visit&& visit) {
    const unsigned short* __restrict__ col = d.nbr16 + i;
    const size_t stride = (size_t)d.nbr_stride;
    const float4* __restrict__ s_pv = win.sv.a[0];
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        const int w0 = __ldg(col + (size_t)k * stride), w1 = __ldg(col + (size_t)(k + 1) * stride);
        const int w2 = __ldg(col + (size_t)(k + 2) * stride), w3 = __ldg(col + (size_t)(k + 3) * stride);
        const float4 p0 = s_pv[w0], p1 = s_pv[w1], p2 = s_pv[w2], p3 = s_pv[w3];
        float3 R;
        R = make_float3(pi.x - p0.x, pi.y - p0.y, pi.z - p0.z); visit(win.sv, w0, p0, R, dist2(R));
        R = make_float3(pi.x - p1.x, pi.y - p1.y, pi.z - p1.z); visit(win.sv, w1, p1, R, dist2(R));
        R = make_float3(pi.x - p2.x, pi.y - p2.y, pi.z - p2.z); visit(win.sv, w2, p2, R, dist2(R));
        R = make_float3(pi.x - p3.x, pi.y - p3.y, pi.z - p3.z); visit(win.sv, w3, p3, R, dist2(R));
    }
    for (; k < n; k++) {
        const int w = __ldg(col + (size_t)k * stride);
        const float4 pj = s_pv[w];
        const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        visit(win.sv, w, pj, R, dist2(R));
    }
}

// Unstaged chunk (window larger than the shared-memory budget): walk the cells in global memory.
template <int NPAY, class Visit>
__device__ __forceinline__ void window_walk_global(const Consts& c, const Dev& d, const Window<NPAY>& win, int i, float4 pi, Visit&& visit) {
    for_all_neighbors(c, d, i, pi, [&](int j, float4 pj, float3 R, float r2) { visit(win.gv, j, pj, R, r2); });
}

// All neighbours of fluid particle i in walk order.  LIST: use the list recorded by the density
// pass when the row fits (count <= kmax) and the chunk is staged; otherwise re-derive them.
template <bool LIST, int NPAY, class Visit>
__device__ __forceinline__ void window_neighbors(const Consts& c, const Dev& d, const Window<NPAY>& win, int i, float4 pi, Visit&& visit) {
    if (!win.staged) {
        window_walk_global(c, d, win, i, pi, visit);
        return;
    }
    if (LIST) {
        const int n = d.nbr_count[i];
        if (n <= d.nbr_kmax) {
            window_list(d, win, i, n, pi, visit);
            return;
        }
    }
    window_walk(c, d, win, i, pi, visit);
}
