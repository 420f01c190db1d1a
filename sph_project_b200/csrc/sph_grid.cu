// sph_grid.cu — uniform-grid spatial hash: cell index + histogram, exclusive scan, deterministic
// counting sort and ping-pong gather of the persistent particle fields.
//
// Replaces BaseContainer.init_grid / PrefixSumExecutor.run / reorder_particles
// (reference base_container.py:495-547).  Differences by design:
//   * flatten is x-fastest / z-slowest (the reference's z-fastest order is not part of its API);
//   * in-cell order is made deterministic (ascending previous index == stable counting sort ==
//     the oracle's order) by a tiny per-cell sort of the scattered permutation, so runs are
//     bit-reproducible although the histogram ranks come from atomics;
//   * fields are gathered once into the other half of a ping-pong pair (156 B/particle of traffic
//     instead of the reference's scatter + copy-back, 304 B/particle);
//   * the gather also flags the bricks (compact tiles of cells, sph_brick.cuh) that own fluid rows; a one-block scan
//     compacts them into the work list of the persistent sweep kernels.
#include "sph_kernels.h"
#include "sph_brick.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// slab: retired particles (ghosts of the previous step, particles that migrated away) are binned
// into a trash cell behind the grid, so the sort itself drops them
__global__ void __launch_bounds__(SPH_BLOCK) k_cell_index(Consts c, Dev d, int slab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    float4 p = d.pv[i];
    int flat = flatten(c, cell_of(c, p.x, p.y, p.z));
    if (slab && d.ghost_slot[i] == 2) flat = c.ncell;
    d.grid_id[i] = flat;
    d.rank[i] = atomicAdd(d.cell_count + flat, 1);
}

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_sums[32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
        int si = warp_inclusive_scan(s);
        warp_sums[lane] = si - s;   // exclusive prefix of the warp sums
        if (lane == 31) block_total = si;
    }
    __syncthreads();
    int res = inc - v + warp_sums[wid];
    *total = block_total;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const int* __restrict__ in, int n, int* __restrict__ tile_sums) {
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
    if (base + SCAN_ITEMS <= n) {
        const int4* p = reinterpret_cast<const int4*>(in + base);
        int4 a = p[0], b = p[1];
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += in[base + k];
    }
    int total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int* tile_sums, int n_tiles) {
    // single block: exclusive scan of the tile sums in place
    int carry = 0;
    for (int base = 0; base < n_tiles; base += 1024) {
        int idx = base + threadIdx.x;
        int v = idx < n_tiles ? tile_sums[idx] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        if (idx < n_tiles) tile_sums[idx] = ex + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const int* __restrict__ in, int n, const int* __restrict__ tile_sums,
                                                            int* __restrict__ out) {
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    // out[n] = grand total
    if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) out[n] = ex;
}

__global__ void __launch_bounds__(SPH_BLOCK) k_scatter_perm(Consts c, Dev d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    d.perm[d.cell_start[d.grid_id[i]] + d.rank[i]] = i;
}

// canonical in-cell order: ascending previous index.  One thread per cell; cells hold ~10
// particles, and the previous order was already sorted, so the insertion sort is nearly a no-op.
__global__ void __launch_bounds__(SPH_BLOCK) k_sort_cells(Consts c, Dev d) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= c.ncell) return;
    const int s = d.cell_start[cell], e = d.cell_start[cell + 1];
    for (int a = s + 1; a < e; a++) {
        int key = d.perm[a];
        int b = a - 1;
        while (b >= s && d.perm[b] > key) {
            d.perm[b + 1] = d.perm[b];
            b--;
        }
        d.perm[b + 1] = key;
    }
}

// Gather of every persistent field into the other half of its ping-pong pair.  Also flags the bricks (sph_brick.cuh)
// that own at least one row a sweep will work on: fluid particles, and emitter particles that are parked as rigid until
// they cross gravitationUpper (base_solver.py:651-677) and may turn fluid before the next sort.
__global__ void __launch_bounds__(SPH_BLOCK) k_gather(Consts c, Dev d, int with_ghost_slot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c.N) return;
    const int i = d.perm[k];
    const float4 pv = d.pv[i];
    const int obj = d.object_id[i];
    const int flat = d.grid_id[i];
    d.pv_alt[k] = pv;
    d.vm_alt[k] = d.vm[i];
    d.x0_alt[3 * k + 0] = d.x0[3 * i + 0];
    d.x0_alt[3 * k + 1] = d.x0[3 * i + 1];
    d.x0_alt[3 * k + 2] = d.x0[3 * i + 2];
    d.rho_alt[k] = d.rho[i];
    d.object_id_alt[k] = obj;
    d.material_alt[k] = d.material[i];
    d.color_alt[3 * k + 0] = d.color[3 * i + 0];
    d.color_alt[3 * k + 1] = d.color[3 * i + 1];
    d.color_alt[3 * k + 2] = d.color[3 * i + 2];
    d.is_dynamic_alt[k] = d.is_dynamic[i];
    d.grid_id_alt[k] = flat;
    d.uid_alt[k] = d.uid[i];
    if (with_ghost_slot) d.ghost_slot_alt[k] = d.ghost_slot[i];
    bool works = pv.w > 0.0f;
    if (!works && obj >= 0 && obj < SPH_MAX_OBJECTS) works = d.object_material[obj] == SPH_MATERIAL_FLUID;
    if (works && SPH_IS_ROW(c, k) && flat < c.ncell) {
        const int cx = flat % c.nx, cy = (flat / c.nx) % c.ny, cz = flat / (c.nx * c.ny);
        d.brick_flag[((cz / BRK_Z) * c.nby + cy / BRK_Y) * c.nbx + cx / BRK_X] = 1;   // same value from every writer
    }
}

// Same flags from the sorted arrays (host edits of materials after the sort)
__global__ void __launch_bounds__(SPH_BLOCK) k_flag_bricks(Consts c, Dev d) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c.N || !SPH_IS_ROW(c, k)) return;
    const int obj = d.object_id[k], flat = d.grid_id[k];
    bool works = d.pv[k].w > 0.0f;
    if (!works && obj >= 0 && obj < SPH_MAX_OBJECTS) works = d.object_material[obj] == SPH_MATERIAL_FLUID;
    if (works && flat < c.ncell) {
        const int cx = flat % c.nx, cy = (flat / c.nx) % c.ny, cz = flat / (c.nx * c.ny);
        d.brick_flag[((cz / BRK_Z) * c.nby + cy / BRK_Y) * c.nbx + cx / BRK_X] = 1;
    }
}

// List of the flagged bricks + the control words of the persistent brick kernels (one block).  Only the brick layers
// that can hold rows of this rank are scanned; bricks whose window reaches into a neighbour rank's layer (ghosts) go
// last, so that the sweeps of a peer loop meet them when the neighbour has long finished the sweep they depend on.
__global__ void __launch_bounds__(1024) k_brick_compact(Consts c, Dev d, int nbricks) {
    constexpr int ITEMS = 8;   // consecutive bricks per thread and trip
    const int per_layer = c.nbx * c.nby;
    const int bz_lo = max(c.z_lo / BRK_Z, 0), bz_hi = min((c.z_hi - 1) / BRK_Z + 1, c.nbz);
    const int first = bz_lo * per_layer, last = min(bz_hi * per_layer, nbricks);
    const int passes = (c.ghost_lo || c.ghost_hi) ? 2 : 1;
    int carry = 0;
    for (int pass = 0; pass < passes; pass++) {
        for (int base = first; base < last; base += 1024 * ITEMS) {
            const int b0 = base + (int)threadIdx.x * ITEMS;
            unsigned mask = 0;
#pragma unroll
            for (int k = 0; k < ITEMS; k++) {
                const int b = b0 + k;
                if (b < last && d.brick_flag[b] != 0) {
                    const int z0 = (b / per_layer) * BRK_Z;   // window layers z0 - 1 .. z0 + BRK_Z
                    const bool edge = (c.ghost_lo && z0 - 1 <= c.z_lo - 1) || (c.ghost_hi && z0 + BRK_Z >= c.z_hi);
                    if ((edge ? 1 : 0) == pass) mask |= 1u << k;
                }
            }
            int total;
            int pos = carry + block_exclusive_scan(__popc(mask), &total);
#pragma unroll
            for (int k = 0; k < ITEMS; k++)
                if (mask & (1u << k)) d.brick_list[pos++] = b0 + k;
            carry += total;
        }
    }
    if (threadIdx.x == 0) {
        d.brick_ctl[BCTL_ACTIVE] = carry;
        d.brick_ctl[BCTL_TICKET] = 0;
        d.brick_ctl[BCTL_FINISHED] = 0;
        d.brick_ctl[BCTL_WMAX_SEEN] = 0;
        d.brick_ctl[BCTL_OVERFLOWS] = 0;
    }
}

}  // namespace

// exclusive scan of in[0..n) into out[0..n], out[n] = total
void sph_exclusive_scan(SphHandle* h, const int* in, int n, int* out) {
    if (n <= 0) {
        cudaMemsetAsync(out, 0, sizeof(int), h->stream);
        return;
    }
    const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    { SphProf p(h, "k_scan_tile_sums"); k_scan_tile_sums<<<tiles, SCAN_THREADS, 0, h->stream>>>(in, n, h->d.scan_tmp); }
    { SphProf p(h, "k_scan_sums"); k_scan_sums<<<1, 1024, 0, h->stream>>>(h->d.scan_tmp, tiles); }
    { SphProf p(h, "k_scan_final"); k_scan_final<<<tiles, SCAN_THREADS, 0, h->stream>>>(in, n, h->d.scan_tmp, out); }
    h->launches += 3;
}

// re-derive the brick work list when materials were edited since the sort
void sph_bricks_refresh(SphHandle* h) {
    if (!h->bricks_dirty) return;
    h->bricks_dirty = false;
    if (!h->sorted_valid) return;   // the next sort rebuilds it
    cudaMemsetAsync(h->d.brick_flag, 0, sizeof(int) * (size_t)h->nbricks, h->stream);
    if (h->c.N > 0) k_flag_bricks<<<(h->c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, h->stream>>>(h->c, h->d);
    k_brick_compact<<<1, 1024, 0, h->stream>>>(h->c, h->d, h->nbricks);
    h->launches += 2;
}

template <class T>
static inline void swap_ptr(T*& a, T*& b) { T* t = a; a = b; b = t; }

int sph_sort_particles(SphHandle* h) {
    Consts& c = h->c;
    Dev& d = h->d;
    cudaStream_t st = h->stream;
    const bool slab = sph_is_slab(h);
    int rc;
    if (slab && (rc = sph_slab_pre_sort(h))) return rc;   // migration + ghost import (appends, retires)
    cudaMemsetAsync(d.cell_count, 0, sizeof(int) * ((size_t)c.ncell + 1), st);
    int nb = (c.N + SPH_BLOCK - 1) / SPH_BLOCK;
    if (c.N > 0) {
        SphProf p(h, "k_cell_index");
        k_cell_index<<<nb, SPH_BLOCK, 0, st>>>(c, d, slab ? 1 : 0);
        h->launches++;
    }
    sph_exclusive_scan(h, d.cell_count, c.ncell + 1, d.cell_start);
    if (c.N > 0) {
        { SphProf p(h, "k_scatter_perm"); k_scatter_perm<<<nb, SPH_BLOCK, 0, st>>>(c, d); }
        h->launches++;
    }
    if (slab) {
        if ((rc = sph_slab_post_scan(h))) return rc;   // live count, owned range, halo ranges (one host sync)
        nb = (c.N + SPH_BLOCK - 1) / SPH_BLOCK;
    } else {
        c.row_begin = 0;
        c.row_end = c.N;
    }
    if (c.N > 0) {
        { SphProf p(h, "k_sort_cells"); k_sort_cells<<<(c.ncell + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, st>>>(c, d); }
        cudaMemsetAsync(d.brick_flag, 0, sizeof(int) * (size_t)h->nbricks, st);
        { SphProf p(h, "k_gather"); k_gather<<<nb, SPH_BLOCK, 0, st>>>(c, d, slab ? 1 : 0); }
        h->launches += 2;
        swap_ptr(d.pv, d.pv_alt);
        swap_ptr(d.vm, d.vm_alt);
        swap_ptr(d.x0, d.x0_alt);
        swap_ptr(d.rho, d.rho_alt);
        swap_ptr(d.object_id, d.object_id_alt);
        swap_ptr(d.material, d.material_alt);
        swap_ptr(d.color, d.color_alt);
        swap_ptr(d.is_dynamic, d.is_dynamic_alt);
        swap_ptr(d.grid_id, d.grid_id_alt);
        swap_ptr(d.uid, d.uid_alt);
        swap_ptr(d.ghost_slot, d.ghost_slot_alt);
    } else {
        cudaMemsetAsync(d.brick_flag, 0, sizeof(int) * (size_t)h->nbricks, st);
    }
    {
        SphProf p(h, "k_brick_compact");
        k_brick_compact<<<1, 1024, 0, st>>>(c, d, h->nbricks);
        h->launches++;
    }
    h->sorted_valid = true;
    h->bricks_dirty = false;
    h->list_valid = false;
    h->ghost_stale = 0;   // the ghosts were just re-imported with their owners' current state
    return cudaGetLastError() == cudaSuccess ? SPH_OK : SPH_E_CUDA;
}
