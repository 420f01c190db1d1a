// sph_window.cuh — shared-memory candidate windows staged by TMA bulk copies (neighbour-list build).
//
// A CTA owns a chunk of SPH_BLOCK consecutive particles of the cell-sorted SoA.  Because the
// flatten is x-fastest, the 27-cell windows of all particles of the chunk are covered by at most 9
// contiguous index ranges of the sorted pv array (one per (dz, dy) row offset, overlapping ranges
// merged).  One elected thread issues cp.async.bulk (TMA, SASS UBLKCP) copies of those ranges
// (float4 per particle, so any range is 16-byte aligned) and the CTA waits on one mbarrier.  The
// ~250 distance tests per particle of the build pass then read shared memory (threads of one cell
// test the same candidate: broadcast, conflict-free) instead of issuing L1 requests.
//
// Measured (profiles/r01_*): staging windows for the list CONSUMERS (payload + 16-bit slot lists)
// was 2x slower than gathering 32-byte records through L1 (low occupancy from 48 KB windows, long
// scoreboard stalls on the list reads), so only the build pass uses the window.
//
// Descriptor per chunk (written once per sort by k_chunk_windows, 40 ints):
//   [0] total window entries   [1] number of merged copies   [2] fluid rows in the chunk   [3] -
//   [4..12]  gstart[r]  first sorted index of row-offset r's range
//   [13..21] sbase[r]   window slot of that first particle
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

struct Window {
    const float4* s_pv;  // staged positions (+-V in w)
    const int* desc;     // shared copy of the chunk descriptor
    bool staged;
};

// Stage the chunk's pv window: call from all threads of the CTA, before any divergence.
__device__ __forceinline__ Window window_open(const Dev& d, float4* smem, int wmax, int* s_desc, unsigned long long* s_mbar) {
    const int tid = threadIdx.x;
    if (tid < SPH_DESC_INTS) s_desc[tid] = d.chunk_desc[(size_t)blockIdx.x * SPH_DESC_INTS + tid];
    if (tid == 0) {
        mbar_init(s_mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    Window w;
    w.desc = s_desc;
    w.s_pv = smem;
    const int total = s_desc[0];
    // chunks without fluid rows (as of the last sort) have nothing to sum: do not stage; a row that
    // turned fluid since (emitter) walks global memory instead
    w.staged = s_desc[2] > 0 && total <= wmax;
    if (w.staged) {
        if (tid == 0) {
            mbar_arrive_expect_tx(s_mbar, (unsigned)total * 16u);
            const int ncp = s_desc[1];
            int s = 0;
            for (int c = 0; c < ncp; c++) {
                const int g = s_desc[22 + c], n = s_desc[31 + c];
                tma_bulk_load(smem + s, d.pv + g, (unsigned)n * 16u, s_mbar);
                s += n;
            }
        }
        mbar_wait(s_mbar, 0);
    }
    return w;
}

// Walk the 27-cell window of particle i with candidates read from the staged window.
// visit(j, pj, R, r2) for every accepted neighbour j (sorted index), same order as for_all_neighbors.
template <class Visit>
__device__ __forceinline__ void window_walk(const Consts& c, const Dev& d, const Window& win, int i, float4 pi, Visit&& visit) {
    const int3 g = cell_of(c, pi.x, pi.y, pi.z);
    const int xlo = max(g.x - 1, 0), xhi = min(g.x + 1, c.nx - 1);
    const float4* __restrict__ s_pv = win.s_pv;
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
                if (r2 < c.h2_thresh && j != i) visit(j, pj, R, r2);
            }
        }
    }
}
