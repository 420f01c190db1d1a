// sph_brick.cuh — compact CTA tiles ("bricks") with TMA-staged shared-memory windows for every neighbour sweep.
//
// A brick is SPH_BRICK_X x SPH_BRICK_Y x SPH_BRICK_Z cells of the uniform grid (default 4 x 4 x 3: 48 cells, ~480
// fluid particles at rest density — a cell of edge h = 4 r holds 10 particles of volume 0.8 (2 r)^3 — i.e. one row
// per thread of a 512-thread CTA with a little head-room).  Its particles and the particles of the one-cell halo around it — the 27-cell
// neighbourhoods of everything the brick owns — are (BY + 2)(BZ + 2) contiguous runs of the cell-sorted arrays,
// because the flatten is x-fastest: run (y, z) covers cells x0-1 .. x0+BX of that grid row.  One CTA works on one
// brick at a time:
//   1. 252 threads read the cell_start entries that delimit the runs (sub-table T, kept in shared memory: it is also
//      the map from a window slot back to a sorted index and the candidate ranges of the list build);
//   2. the runs of up to three float4 arrays (pv and the sweep's payloads) are copied into shared memory by TMA bulk
//      copies (cp.async.bulk + mbarrier; SASS UBLKCP), one copy per run and array, issued by the first 36 threads;
//   3. every owned fluid particle streams its neighbour list — 16-bit WINDOW SLOTS, 16 per 256-bit load — and reads
//      neighbour j's position and payload from shared memory (LDS.128) instead of gathering it through the L1.
// The halo makes the window (BX+2)(BY+2)(BZ+2) / (BX BY BZ) = 3.75x the owned data (the round-1 chunks of 128
// consecutive particles were 1 x 1 x 13-cell sticks: ~10x).  CTAs are persistent: they draw bricks from a ticket
// counter over the compacted list of bricks that own fluid rows (built by the sort), so boundary-only and empty
// bricks cost nothing and the tail is balanced dynamically.
//
// Load balance inside a CTA (measured on 4 x 4 x 4 bricks = 640 rows per 512 threads: the four warps with a second
// pass kept the other twelve at the end-of-brick barrier, 28 % of all warp time): bricks are sized to at most one row
// per thread, the list build leaves a stably compacted list of the brick's working rows (no lanes parked on boundary
// particles), and warps draw groups of 32 consecutive rows of it from a shared counter.
//
// Rows without a usable list (more than nbr_kmax neighbours, a window beyond 16 bits, a particle that left its sorted
// cell since the sort) and windows that exceed the shared-memory budget fall back to global memory with the same
// visiting order, so every path sums in the order of the reference's for_all_neighbors (base_container.py:549-560)
// on this library's x-fastest grid.
#pragma once

#include "sph_common.cuh"
#include "sph_kernels.h"

#ifndef SPH_BRICK_X
#define SPH_BRICK_X 4
#endif
#ifndef SPH_BRICK_Y
#define SPH_BRICK_Y 4
#endif
#ifndef SPH_BRICK_Z
#define SPH_BRICK_Z 3
#endif
#ifndef SPH_BRICK_THREADS
#define SPH_BRICK_THREADS 512
#endif

constexpr int BRK_X = SPH_BRICK_X, BRK_Y = SPH_BRICK_Y, BRK_Z = SPH_BRICK_Z;
constexpr int BRK_RY = BRK_Y + 2, BRK_RZ = BRK_Z + 2;
constexpr int BRK_RUNS = BRK_RY * BRK_RZ;        // window runs (one per (y, z) row of the haloed brick)
constexpr int BRK_OWN_RUNS = BRK_Y * BRK_Z;      // runs that hold owned particles
constexpr int BRK_TW = BRK_X + 3;                // cell_start entries per run: cells x0-1 .. x0+BX and the end
constexpr int BRK_CELLS = BRK_X * BRK_Y * BRK_Z;
constexpr int BRK_WARPS = SPH_BRICK_THREADS / 32;
constexpr int BRK_SORT_PASSES = 4;                                   // rows per thread the build can compact
constexpr int BRK_ROWS_MAX = BRK_SORT_PASSES * SPH_BRICK_THREADS;   // owned particles per brick beyond which the row list is not built
static_assert(BRK_RUNS <= SPH_BRICK_THREADS, "one thread per run issues the TMA copies");

enum BrickCtl { BCTL_ACTIVE = 0, BCTL_TICKET = 1, BCTL_FINISHED = 2, BCTL_WMAX_SEEN = 3, BCTL_OVERFLOWS = 4, BCTL_COUNT = 8 };

// ---- mbarrier / TMA bulk copy primitives ---------------------------------------------------------
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
__device__ __forceinline__ void tma_bulk_load(unsigned smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 16 list entries (16-bit window slots) in one 256-bit load.  NC: the list was written by an earlier launch — read-only
// path, streamed past the L1 and marked evict-first in the L2 so that it does not displace the particle arrays the
// windows are staged from.
template <bool NC>
__device__ __forceinline__ void ld_slots16(const unsigned short* p, unsigned (&w)[8]) {
    if (NC) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(p));
    } else {
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(p)
                     : "memory");
    }
}

// shared-memory loads by 32-bit shared address (keeps generic-pointer conversions out of the inner loops)
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds64(unsigned addr) {   // first two components only
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ int brk_ld_volatile_i32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 brk_ld_volatile_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- per-CTA brick state in shared memory -----------------------------------------------------------
// Tables of one brick.  Two sets per CTA: while the CTA works on one brick, warp 0 prepares the tables of the next,
// so a single barrier and the TMA round trip are all that separates two bricks.
struct BrickShared {
    int T[BRK_RUNS][BRK_TW];     // T[r][k] = cell_start of cell x0-1+k in run r (sorted indices); all 0 for rows outside the grid
    int S[BRK_RUNS + 1];         // window slot of the first particle of run r (prefix of the run lengths)
    int OP[BRK_OWN_RUNS + 1];    // prefix of the owned-run lengths: flat owned index -> run
    int brick;                   // brick id, -1 when the ticket counter ran out
    int ordinal;                 // its position in the active list (index of brick_nf)
    int nf;                      // working rows in the compacted row list (row_order), -1: none, walk the owned rows as they come
    int group;                   // next group of 32 listed rows to hand out
    int staged;                  // window is in shared memory (else: same slots, fetched from global memory)
    int x0, y0, z0;              // first owned cell
    int ghost_slots;             // window slots in a neighbour rank's layers (their payload may come from its memory)
};
// fused ghost reads of a Z-slab peer loop: where the neighbours' payload lives and what to wait for
struct BrickGhost {
    const float4* arr[2];
    const int* counter[2];
    const int* layout[2];        // the neighbours' (own_begin, own_end, ...) of the current sort
    int delta[2];                // ghost sorted index + delta = the neighbour's sorted index of the same particle (set with `ready`)
    int want;                    // completed writers the neighbours must have counted
    int ready[2];
    int own_begin, own_end;
};
struct BrickSmem {
    unsigned long long mbar;
    BrickShared tab[2];
    BrickGhost gh;
};

struct Brick {
    BrickSmem* smem;
    BrickShared* sh;             // tables of the brick in hand
    unsigned a0, a1, a2;         // shared addresses of the staged arrays (a0 = pv window), 16 B per slot
    const float4* g0;            // their global sources
    const float4* g1;
    const float4* g2;
    unsigned phase;              // mbarrier parity of the next wait
    int parity;                  // which table set is in hand
    int wmax;
    bool stage, sorted;
    bool fused;                  // payload array 1 of ghost particles is read from the neighbour ranks' memory
    int b_pre, o_pre, nf_pre, t_pre;   // thread 0: prefetched brick id, ordinal, listed-row count (brick after next) and ticket (the one after)
};

// one owned particle: sorted index and window slot
struct BrickRow {
    int i;
    int slot;
    int run;   // window run (y, z row of the haloed brick) the particle lives in
};

// reference to neighbour j as the visitor sees it: a window slot (list paths) or a sorted index (walk path)
struct NbrRef {
    int v;
    bool is_slot;
};

__device__ __forceinline__ int brk_warp_inclusive_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)(threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// window slot -> sorted index (rare consumers: rigid wrench accumulation; the unstaged list path)
__device__ __forceinline__ int brick_slot_to_index(const BrickShared& sh, int slot) {
    int r = 0;
#pragma unroll 1
    while (r + 1 < BRK_RUNS && slot >= sh.S[r + 1]) r++;
    return sh.T[r][0] + (slot - sh.S[r]);
}
__device__ __forceinline__ int brick_nbr_index(const Brick& bk, NbrRef ref) { return ref.is_slot ? brick_slot_to_index(*bk.sh, ref.v) : ref.v; }

// thread 0: brick id / ordinal / listed rows of ticket `ticket` (loads issued now, consumed one brick later)
__device__ __forceinline__ void brick_prefetch(const Dev& d, Brick& bk, int ticket) {
    const bool live = ticket < d.brick_ctl[BCTL_ACTIVE];
    bk.o_pre = ticket;
    bk.b_pre = live ? __ldg(d.brick_list + ticket) : -1;
    bk.nf_pre = (live && bk.sorted) ? d.brick_nf[ticket] : -1;
}

// warp 0: tables of the brick thread 0 has prefetched, into `tab`; thread 0 then prefetches the one after
__device__ __forceinline__ void brick_prepare(const Consts& c, const Dev& d, Brick& bk, BrickShared& tab) {
    const int lane = threadIdx.x;   // warp 0
    const int b = __shfl_sync(0xffffffffu, bk.b_pre, 0);
    const int ordinal = __shfl_sync(0xffffffffu, bk.o_pre, 0);
    const int nf = __shfl_sync(0xffffffffu, bk.nf_pre, 0);
    if (lane == 0) {
        tab.brick = b;
        tab.ordinal = ordinal;
        tab.nf = nf;
        tab.group = 0;
        if (b >= 0) {
            brick_prefetch(d, bk, bk.t_pre);
            bk.t_pre = atomicAdd(d.brick_ctl + BCTL_TICKET, 1);
        }
    }
    if (b >= 0) {
        const int bx = b % c.nbx, by = (b / c.nbx) % c.nby, bz = b / (c.nbx * c.nby);
        const int x0 = bx * BRK_X, y0 = by * BRK_Y, z0 = bz * BRK_Z;
#pragma unroll 1
        for (int e = lane; e < BRK_RUNS * BRK_TW; e += 32) {
            const int r = e / BRK_TW, k = e - r * BRK_TW;
            const int y = y0 - 1 + r % BRK_RY, z = z0 - 1 + r / BRK_RY;
            int v = 0;
            if (y >= 0 && y < c.ny && z >= 0 && z < c.nz) {
                const int x = min(max(x0 - 1 + k, 0), c.nx);   // x == nx: end of the grid row
                v = __ldg(d.cell_start + ((z * c.ny + y) * c.nx + x));
            }
            tab.T[r][k] = v;
        }
        __syncwarp();
        int carry = 0;   // prefix sums of the run lengths (window slots) and of the owned-run lengths
        int ghost = 0;   // slots of runs in a neighbour rank's layer
#pragma unroll
        for (int r0 = 0; r0 < BRK_RUNS; r0 += 32) {
            const int r = r0 + lane;
            const int len = r < BRK_RUNS ? tab.T[r][BRK_TW - 1] - tab.T[r][0] : 0;
            const int inc = brk_warp_inclusive_scan(len);
            if (r < BRK_RUNS) tab.S[r] = carry + inc - len;
            carry += __shfl_sync(0xffffffffu, inc, 31);
            const int z = z0 - 1 + r / BRK_RY;
            if (r < BRK_RUNS && ((c.ghost_lo && z == c.z_lo - 1) || (c.ghost_hi && z == c.z_hi))) ghost += len;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ghost += __shfl_xor_sync(0xffffffffu, ghost, o);
        int ocarry = 0;
#pragma unroll
        for (int q0 = 0; q0 < BRK_OWN_RUNS; q0 += 32) {
            const int q = q0 + lane;
            int len = 0;
            if (q < BRK_OWN_RUNS) {
                const int r = (q / BRK_Y + 1) * BRK_RY + (q % BRK_Y + 1);
                len = tab.T[r][BRK_X + 1] - tab.T[r][1];
            }
            const int inc = brk_warp_inclusive_scan(len);
            if (q < BRK_OWN_RUNS) tab.OP[q] = ocarry + inc - len;
            ocarry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            tab.S[BRK_RUNS] = carry;
            tab.OP[BRK_OWN_RUNS] = ocarry;
            tab.staged = bk.stage && carry <= bk.wmax;
            tab.x0 = x0; tab.y0 = y0; tab.z0 = z0;
            tab.ghost_slots = ghost;
            if (carry > d.brick_ctl[BCTL_WMAX_SEEN]) atomicMax(d.brick_ctl + BCTL_WMAX_SEEN, carry);
            if (bk.stage && carry > bk.wmax) atomicAdd(d.brick_ctl + BCTL_OVERFLOWS, 1);
        }
    }
    __syncwarp();
}

// Wait (one lane) until the neighbour on `side` has completed the sweep whose output this one reads, then fix the index
// shift between my ghost rows and its owned rows: my lower ghosts are the lower neighbour's last owned rows, my upper
// ones the upper neighbour's first (its layout words were published by its sort, long before that sweep).
__device__ __forceinline__ void brick_ghost_wait(BrickGhost& gh, int side) {
    if (*(volatile int*)&gh.ready[side] != 0) return;
    while (brk_ld_volatile_i32(gh.counter[side]) < gh.want) __nanosleep(64);
    __threadfence_system();
    const int d = side == 0 ? brk_ld_volatile_i32(gh.layout[0] + 1) - gh.own_begin : brk_ld_volatile_i32(gh.layout[1] + 0) - gh.own_end;
    *(volatile int*)&gh.delta[side] = d;   // several warps may get here at once: they all store the same value
    __threadfence_block();
    *(volatile int*)&gh.ready[side] = 1;
}

// CTA-wide, once per kernel: barrier, first tickets, tables of the first brick.
// stage: copy windows into shared memory; sorted: walk the rows through the list build's row lists.
__device__ __forceinline__ void brick_begin(const Consts& c, const Dev& d, Brick& bk, BrickSmem* smem, float4* window, int wmax,
                                            const float4* g0, const float4* g1, const float4* g2, bool stage, bool sorted,
                                            const PeerLinks* peer = nullptr) {
    bk.smem = smem;
    bk.a0 = smem_u32(window); bk.a1 = bk.a0 + 16u * (unsigned)wmax; bk.a2 = bk.a1 + 16u * (unsigned)wmax;
    bk.g0 = g0; bk.g1 = g1; bk.g2 = g2;
    bk.phase = 0;
    bk.parity = 0;
    bk.wmax = wmax;
    bk.stage = stage; bk.sorted = sorted;
    bk.b_pre = -1; bk.o_pre = 0; bk.nf_pre = -1; bk.t_pre = 0;
    bk.sh = &smem->tab[0];
    bk.fused = peer != nullptr && peer->fused != 0;
    if (threadIdx.x == 0) {
        if (bk.fused) {   // where the neighbours keep my ghosts' payload, and how far they must have counted
            BrickGhost& gh = smem->gh;
            for (int side = 0; side < 2; side++) {
                gh.arr[side] = peer->ghost_arr[side];
                gh.counter[side] = peer->ghost_counter[side];
                gh.layout[side] = peer->ghost_layout[side];
                gh.ready[side] = 0;
                gh.delta[side] = 0;
            }
            gh.want = *(volatile const int*)peer->my_counter;
            gh.own_begin = peer->own_begin;
            gh.own_end = peer->own_end;
        }
        mbar_init(&smem->mbar, 1);
        fence_mbar_init();
        brick_prefetch(d, bk, atomicAdd(d.brick_ctl + BCTL_TICKET, 1));
        bk.t_pre = atomicAdd(d.brick_ctl + BCTL_TICKET, 1);
    }
    if (threadIdx.x < 32) brick_prepare(c, d, bk, smem->tab[0]);
    __syncthreads();
}

// CTA-wide: false when the bricks have run out; else the window of the brick in hand is staged (NARR arrays by TMA
// bulk copies, budget permitting) and warp 0 has prepared the tables of the next brick.
template <int NARR>
__device__ __forceinline__ bool brick_stage(const Consts& c, const Dev& d, Brick& bk) {
    BrickShared& sh = *bk.sh;
    if (sh.brick < 0) return false;
    const int tid = threadIdx.x;
    const bool ghosts = bk.fused && NARR > 1 && sh.ghost_slots > 0;   // uniform over the CTA
    if (sh.staged) {
        // lane 0 of every warp issues the copies of runs warp, warp + BRK_WARPS, ...: the issue is spread over the four
        // schedulers instead of being serialised inside one warp (UBLKCP takes its operands from uniform registers)
        const int lane = tid & 31;
        for (int r = tid >> 5; r < BRK_RUNS; r += BRK_WARPS) {
            const int g = sh.T[r][0], n = sh.T[r][BRK_TW - 1] - g, s = sh.S[r];
            if (n <= 0) continue;
            int side = -1;   // run of a neighbour rank's layer whose payload is read from that rank's memory
            if (ghosts) {
                const int z = sh.z0 - 1 + r / BRK_RY;
                side = (c.ghost_lo && z == c.z_lo - 1) ? 0 : ((c.ghost_hi && z == c.z_hi) ? 1 : -1);
            }
            if (lane == 0) {
                tma_bulk_load(bk.a0 + 16u * s, bk.g0 + g, (unsigned)n * 16u, &bk.smem->mbar);
                if (NARR > 1 && side < 0) tma_bulk_load(bk.a1 + 16u * s, bk.g1 + g, (unsigned)n * 16u, &bk.smem->mbar);
                if (NARR > 2) tma_bulk_load(bk.a2 + 16u * s, bk.g2 + g, (unsigned)n * 16u, &bk.smem->mbar);
            }
            if (side >= 0) {
                BrickGhost& gh = bk.smem->gh;
                if (lane == 0) brick_ghost_wait(gh, side);
                __syncwarp();
                const float4* src = gh.arr[side] + g + *(volatile int*)&gh.delta[side];
                for (int e = lane; e < n; e += 32) sts128(bk.a1 + 16u * (unsigned)(s + e), brk_ld_volatile_f4(src + e));
            }
        }
        // the phase cannot complete before this arrival, whatever the copies have already delivered
        if (tid == 0) mbar_arrive_expect_tx(&bk.smem->mbar, ((unsigned)sh.S[BRK_RUNS] * NARR - (ghosts ? (unsigned)sh.ghost_slots : 0u)) * 16u);
    } else if (ghosts) {   // unstaged window: its rows fetch ghost payloads from the neighbours one by one — after their sweep
        if (tid == 0) {
            BrickGhost& gh = bk.smem->gh;
            for (int side = 0; side < 2; side++)
                if (gh.arr[side]) brick_ghost_wait(gh, side);
        }
    }
    if (tid < 32) brick_prepare(c, d, bk, bk.smem->tab[bk.parity ^ 1]);   // overlaps the copies
    if (sh.staged) {
        mbar_wait(&bk.smem->mbar, bk.phase);
        bk.phase ^= 1u;
    }
    if (ghosts) __syncthreads();   // ghost runs were stored by ordinary instructions of several warps / the wait above
    return true;
}

// CTA-wide: done with the brick in hand (its window and tables may be overwritten); the next one is in the other set
__device__ __forceinline__ void brick_advance(Brick& bk) {
    __syncthreads();
    bk.parity ^= 1;
    bk.sh = &bk.smem->tab[bk.parity];
}

// CTA-wide, once at the end of the kernel: the last CTA to finish re-arms the ticket counter for the next launch.
// Returns true in thread 0 of that last CTA (it may then run a grid-level epilogue: everything the other CTAs wrote
// before their own brick_finish is visible to it).
__device__ __forceinline__ bool brick_finish(const Dev& d) {
    bool last = false;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int f = atomicAdd(d.brick_ctl + BCTL_FINISHED, 1);
        if (f == (int)gridDim.x - 1) {
            d.brick_ctl[BCTL_TICKET] = 0;
            d.brick_ctl[BCTL_FINISHED] = 0;
            __threadfence();
            last = true;
        }
    }
    return last;
}

__device__ __forceinline__ int brick_own_count(const Brick& bk) { return bk.sh->OP[BRK_OWN_RUNS]; }

// flat owned index t -> sorted particle index + window slot (branch-free: broadcast reads of the 16-entry prefix)
__device__ __forceinline__ BrickRow brick_own_row(const Brick& bk, int t) {
    const BrickShared& sh = *bk.sh;
    int q = 0;
#pragma unroll
    for (int k = 1; k < BRK_OWN_RUNS; k++) q += (t >= sh.OP[k]) ? 1 : 0;   // t < OP[BRK_OWN_RUNS] is the caller's business
    const int r = (q / BRK_Y + 1) * BRK_RY + (q % BRK_Y + 1);
    const int off = t - sh.OP[q];
    BrickRow row;
    row.i = sh.T[r][1] + off;
    row.slot = sh.S[r] + (sh.T[r][1] - sh.T[r][0]) + off;
    row.run = r;
    return row;
}

// entry `row` of staged array `which` (0: pv) — the owned particle's own data comes from the window as well
__device__ __forceinline__ float4 brick_own_load(const Brick& bk, BrickRow row, int which) {
    if (bk.sh->staged) return lds128((which == 0 ? bk.a0 : (which == 1 ? bk.a1 : bk.a2)) + 16u * (unsigned)row.slot);
    return (which == 0 ? bk.g0 : (which == 1 ? bk.g1 : bk.g2))[row.i];
}

// payload array 1 of sorted index j from global memory: a ghost's comes from its owner rank inside a fused peer loop
__device__ __forceinline__ float4 brick_payload1(const Brick& bk, int j) {
    if (bk.fused) {
        const BrickGhost& gh = bk.smem->gh;
        if (j < gh.own_begin && gh.arr[0]) return brk_ld_volatile_f4(gh.arr[0] + j + *(volatile const int*)&gh.delta[0]);
        if (j >= gh.own_end && gh.arr[1]) return brk_ld_volatile_f4(gh.arr[1] + j + *(volatile const int*)&gh.delta[1]);
    }
    return __ldg(bk.g1 + j);
}

// Neighbour-list rows: nbr_kmax 16-bit words per particle, word 0 = number of neighbours n (SPH_ROW_NO_LIST: the row
// has no list), words 1 .. n = their window slots in walk order.
#define SPH_ROW_NO_LIST 0xffffu

// All neighbours j of owned particle i in walk order: visit(ref, pj, aj, bj, R, r2), aj / bj = entry j of the second /
// third staged array (pv_j itself when the sweep stages fewer).  Returns the number of neighbours.
// A_HALF: the visitor reads only the first two components of aj (64-bit shared loads).
template <bool LIST, bool NC, int NARR, bool A_HALF = false, class Visit>
__device__ __forceinline__ int brick_neighbors(const Consts& c, const Dev& d, const Brick& bk, int i, float4 pi, Visit&& visit) {
    if (LIST) {
        const unsigned short* row = d.nbr16 + (size_t)i * d.nbr_kmax;
        const BrickShared& sh = *bk.sh;
        unsigned w[8];
        ld_slots16<NC>(row, w);   // the count and the first 15 neighbours in one 256-bit load
        const int n = (int)(w[0] & 0xffffu);
        if (n != (int)SPH_ROW_NO_LIST) {
            if (sh.staged) {
                // the window addresses live in registers for the whole row: without the opaque move the compiler
                // re-derives them from the shared-window base (uniform-datapath instructions) for every neighbour
                unsigned a0 = bk.a0, a1 = bk.a1, a2 = bk.a2;
                asm volatile("mov.u32 %0, %0;" : "+r"(a0));
                if (NARR > 1) asm volatile("mov.u32 %0, %0;" : "+r"(a1));
                if (NARR > 2) asm volatile("mov.u32 %0, %0;" : "+r"(a2));
                // One 256-bit load per 16 list words, one neighbour at a time.  Issuing the shared-memory loads of 2-4
                // neighbours together, or one neighbour ahead, was measured slower (batches of 4 need 64 registers, i.e.
                // 2 resident CTAs per SM instead of 3: 186 / 202 us per correction / density-change sweep against
                // 168 / 181 us; profiles/r02_brick_experiments.md).
                int base = 0;
#pragma unroll 1
                for (;;) {
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        if ((u > 0 || base > 0) && base + u <= n) {
                            const unsigned off = (u & 1) ? ((w[u >> 1] >> 12) & 0xffff0u) : ((w[u >> 1] << 4) & 0xffff0u);   // 16 x slot
                            const float4 pj = lds128(a0 + off);
                            const float4 aj = NARR > 1 ? (A_HALF ? lds64(a1 + off) : lds128(a1 + off)) : pj;
                            const float4 bj = NARR > 2 ? lds128(a2 + off) : pj;
                            const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                            visit(NbrRef{(int)(off >> 4), true}, pj, aj, bj, R, dist2(R));
                        }
                    }
                    base += 16;
                    if (base > n) break;
                    ld_slots16<NC>(row + base, w);
                }
            } else {   // window above the shared-memory budget: same list, entries fetched from global memory
#pragma unroll 1
                for (int k = 1; k <= n; k++) {
                    const int slot = NC ? (int)__ldg(row + k) : (int)((const volatile unsigned short*)row)[k];
                    const int j = brick_slot_to_index(sh, slot);
                    const float4 pj = __ldg(bk.g0 + j);
                    const float4 aj = NARR > 1 ? brick_payload1(bk, j) : pj;
                    const float4 bj = NARR > 2 ? __ldg(bk.g2 + j) : pj;
                    const float3 R = make_float3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                    visit(NbrRef{slot, true}, pj, aj, bj, R, dist2(R));
                }
            }
            return n;
        }
    }
    int n = 0;
    for_all_neighbors(c, d, i, pi, [&](int j, float4 pj, float3 R, float r2) {
        const float4 aj = NARR > 1 ? brick_payload1(bk, j) : pj;
        const float4 bj = NARR > 2 ? __ldg(bk.g2 + j) : pj;
        visit(NbrRef{j, false}, pj, aj, bj, R, r2);
        n++;
    });
    return n;
}
