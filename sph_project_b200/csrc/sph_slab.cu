// sph_slab.cu — Z-slab domain decomposition across the GPUs of one NVSwitch box;
// no counterpart in the reference, which is single-device.
//
// One process / handle per GPU.  Rank r owns the cell layers cz in [z_lo, z_hi) of the global grid
// and keeps read-only copies ("ghosts") of the neighbouring ranks' boundary layers z_lo-1 and z_hi
// (support radius = cell size, so one layer suffices for every sweep).  Because the flatten is
// z-slowest and the sort is stable, after every sort
//   [ghost layer z_lo-1][owned layers ...][ghost layer z_hi]
// are three contiguous index ranges, and my boundary layer z_lo (z_hi-1) holds exactly the particles
// of the lower (upper) neighbour's ghost layer, in the same order.  A halo refresh of any field is
// therefore a plain contiguous ncclSend/ncclRecv of an index range — no pack kernels, no index maps.
//
// Once per step, before the sort, particles that left the slab migrate to the neighbour and the
// ghost layers are re-imported (full particle records); inside the solver loops only the fields a
// sweep's inputs changed are refreshed: vm (velocities) and aux (kappa), 16 bytes per ghost each, once per DFSPH
// iteration, plus one ncclAllReduce of the error sum.  All NCCL calls are enqueued on the handle's stream.
//
// Every decision that leads to a collective or point-to-point call is the same on all ranks: the kernels a step
// launches depend on the solver, on iteration counts (all-reduced), and on two flags — "some rank holds dynamic rigid
// particles" and "some rank's boundary volumes are out of date" — that are summed over the ranks at every sort.
//
// NCCL is resolved at run time from the process (torch has already loaded libnccl.so.2); the
// communicator is created from a unique id that the Python host broadcasts with torch.distributed.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sph_kernels.h"

// ---- minimal NCCL ABI (stable across NCCL 2.x) ---------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };
struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId*);
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
    bool ok = false;
};
static NcclApi g_nccl;

static bool load_nccl() {
    if (g_nccl.ok) return true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
#define SYM(field, name)                                                   \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                           \
    if (!g_nccl.field) return false
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.ok = true;
    return true;
}

#define REC_WORDS SPH_SLAB_RECORD_WORDS   // 24 x 4 B per migrating / ghost particle

struct SlabState {
    int rank = 0, world = 1;
    int z_lo = 0, z_hi = 0;
    ncclComm_t comm = nullptr;
    // layout after the last sort
    int own_begin = 0, own_end = 0;       // owned index range
    int send_lo_n = 0, send_hi_n = 0;     // my boundary layers (== neighbours' ghost layers)
    int ghost_lo_n = 0, ghost_hi_n = 0;
    // exchange scratch
    int* flags = nullptr;                 // cap ints
    int* scan = nullptr;                  // cap + 1 ints
    float* sendbuf[2] = {nullptr, nullptr};
    float* recvbuf[2] = {nullptr, nullptr};
    int buf_records = 0;
    int* d_counts = nullptr;              // [0..1] send counts lo/hi, [2..3] recv counts lo/hi
    int* h_ints = nullptr;                // pinned
    int64_t halo_bytes = 0, halo_calls = 0;
    // re-balancing: every `rebalance_period` sorts a boundary may move by one cell layer towards the busier rank
    int sorts = 0, rebalance_period = 0, rebalances = 0;
    int* d_work = nullptr;                // [0..7] my work figures, [8..15] lower neighbour's, [16..23] upper neighbour's
    // peer memory (CUDA IPC over NVLink), optional
    PeerCtl* ctl = nullptr;                          // my control block (device)
    PeerCtl* peer_ctl[SPH_PEER_MAX_RANKS] = {nullptr};   // every rank's block as mapped here (own included)
    PeerCtl** d_peer_ctl = nullptr;                  // the same table in device memory
    float4* peer_vm[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side lo / hi][exporter's vm / vm_alt]
    float4* peer_aux[2] = {nullptr, nullptr};
    float4* export_vm = nullptr;                     // my d.vm at export time (tells which half is current later)
    int imported = 0;
    bool peers_ready = false;
    int64_t peer_pulls = 0;
};

namespace {

int fail(SphHandle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}
int nccl_check(SphHandle* h, int rc, const char* what) {
    if (rc == ncclSuccess) return SPH_OK;
    h->err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error");
    return SPH_E_CUDA;
}
#define NCCL_TRY(h, expr)                           \
    do {                                            \
        int _rc = nccl_check((h), (expr), #expr);   \
        if (_rc) return _rc;                        \
    } while (0)
#define CU_TRY(h, expr)                                                         \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);      \
            return SPH_E_CUDA;                                                  \
        }                                                                       \
    } while (0)

// ghost_slot values
enum { SLOT_OWNED = 0, SLOT_GHOST = 1, SLOT_DEAD = 2 };

// mode 0 (migration): flag live owned particles whose cell layer left [z_lo, z_hi) towards `side`,
//                     and retire every ghost and every leaver (SLOT_DEAD, dropped by the next sort)
// mode 1 (ghost export): flag live owned particles of my boundary layer on `side`
__global__ void __launch_bounds__(SPH_BLOCK) k_slab_flags(Consts c, Dev d, int mode, int side, int z_lo, int z_hi, int* flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    const int slot = d.ghost_slot[i];
    const float4 p = d.pv[i];
    const int cz = cell_of(c, p.x, p.y, p.z).z;
    int f = 0;
    if (mode == 0) {
        if (slot == SLOT_OWNED) f = side == 0 ? (cz < z_lo) : (cz >= z_hi);
    } else {
        if (slot == SLOT_OWNED) f = side == 0 ? (cz == z_lo) : (cz == z_hi - 1);
    }
    flags[i] = f;
}
__global__ void __launch_bounds__(SPH_BLOCK) k_slab_retire(Consts c, Dev d, int z_lo, int z_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    const int slot = d.ghost_slot[i];
    if (slot == SLOT_GHOST) { d.ghost_slot[i] = SLOT_DEAD; return; }
    if (slot != SLOT_OWNED) return;
    const float4 p = d.pv[i];
    const int cz = cell_of(c, p.x, p.y, p.z).z;
    if (cz < z_lo || cz >= z_hi) d.ghost_slot[i] = SLOT_DEAD;
}

// full particle record: pv(4) vm(4) x0(3) rho(1) | object_id material is_dynamic uid color(3) pad
__global__ void __launch_bounds__(SPH_BLOCK) k_slab_pack(Consts c, Dev d, const int* flags, const int* scan, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N || !flags[i]) return;
    float* r = out + (size_t)scan[i] * REC_WORDS;
    int* ri = reinterpret_cast<int*>(r);
    const float4 p = d.pv[i], v = d.vm[i];
    r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = p.w;
    r[4] = v.x; r[5] = v.y; r[6] = v.z; r[7] = v.w;
    r[8] = d.x0[3 * i]; r[9] = d.x0[3 * i + 1]; r[10] = d.x0[3 * i + 2];
    r[11] = d.rho[i];
    ri[12] = d.object_id[i]; ri[13] = d.material[i]; ri[14] = d.is_dynamic[i]; ri[15] = d.uid[i];
    ri[16] = d.color[3 * i]; ri[17] = d.color[3 * i + 1]; ri[18] = d.color[3 * i + 2];
    ri[19] = 0; ri[20] = 0; ri[21] = 0; ri[22] = 0; ri[23] = 0;
}
__global__ void __launch_bounds__(SPH_BLOCK) k_slab_unpack(Dev d, const float* in, int n, int base, int slot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float* r = in + (size_t)k * REC_WORDS;
    const int* ri = reinterpret_cast<const int*>(r);
    const int i = base + k;
    d.pv[i] = make_float4(r[0], r[1], r[2], r[3]);
    d.vm[i] = make_float4(r[4], r[5], r[6], r[7]);
    d.x0[3 * i] = r[8]; d.x0[3 * i + 1] = r[9]; d.x0[3 * i + 2] = r[10];
    d.rho[i] = r[11];
    d.object_id[i] = ri[12]; d.material[i] = ri[13]; d.is_dynamic[i] = ri[14]; d.uid[i] = ri[15];
    d.color[3 * i] = ri[16]; d.color[3 * i + 1] = ri[17]; d.color[3 * i + 2] = ri[18];
    d.ghost_slot[i] = slot;
}

// Work figures of a slab for the re-balancing: a fluid row costs ~20 boundary rows (the sweeps only work on fluid).
// out[0] work of all owned particles, [1] / [2] work of the bottom / top owned layer, [3] / [4] their particle counts
constexpr int SLAB_FLUID_WEIGHT = 20;
__global__ void __launch_bounds__(SPH_BLOCK) k_slab_work(Consts c, Dev d, int z_lo, int z_hi, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int w = 0, lo = 0, hi = 0;
    if (i < c.N && d.ghost_slot[i] == SLOT_OWNED) {
        const float4 p = d.pv[i];
        const int cz = cell_of(c, p.x, p.y, p.z).z;
        if (cz >= z_lo && cz < z_hi) {
            w = p.w > 0.0f ? SLAB_FLUID_WEIGHT : 1;
            lo = cz == z_lo;
            hi = cz == z_hi - 1;
        }
    }
    const unsigned full = 0xffffffffu;
    const int w_all = __reduce_add_sync(full, w), w_lo = __reduce_add_sync(full, lo ? w : 0), w_hi = __reduce_add_sync(full, hi ? w : 0);
    const int n_lo = __reduce_add_sync(full, lo), n_hi = __reduce_add_sync(full, hi);
    if ((threadIdx.x & 31) == 0) {
        if (w_all) atomicAdd(out + 0, w_all);
        if (w_lo) atomicAdd(out + 1, w_lo);
        if (w_hi) atomicAdd(out + 2, w_hi);
        if (n_lo) atomicAdd(out + 3, n_lo);
        if (n_hi) atomicAdd(out + 4, n_hi);
    }
}

int neighbour(const SlabState* s, int side) {
    const int r = side == 0 ? s->rank - 1 : s->rank + 1;
    return (r < 0 || r >= s->world) ? -1 : r;
}

// one phase of the pre-sort exchange: pack flagged particles per side, swap counts, swap records,
// append what arrived with ghost_slot = slot
int exchange_phase(SphHandle* h, int mode, int slot) {
    SlabState* s = h->slab;
    Consts& c = h->c;
    const int nb = (c.N + SPH_BLOCK - 1) / SPH_BLOCK;
    cudaStream_t st = h->stream;
    CU_TRY(h, cudaMemsetAsync(s->d_counts, 0, 4 * sizeof(int), st));
    for (int side = 0; side < 2; side++) {
        if (neighbour(s, side) < 0 || c.N == 0) continue;
        k_slab_flags<<<nb, SPH_BLOCK, 0, st>>>(c, h->d, mode, side, s->z_lo, s->z_hi, s->flags);
        sph_exclusive_scan(h, s->flags, c.N, s->scan);
        k_slab_pack<<<nb, SPH_BLOCK, 0, st>>>(c, h->d, s->flags, s->scan, s->sendbuf[side]);
        CU_TRY(h, cudaMemcpyAsync(s->d_counts + side, s->scan + c.N, sizeof(int), cudaMemcpyDeviceToDevice, st));
        h->launches += 2;
    }
    // counts: device -> neighbours, then one host read of (send, recv) counts
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int side = 0; side < 2; side++) {
        const int nbr = neighbour(s, side);
        if (nbr < 0) continue;
        NCCL_TRY(h, g_nccl.Send(s->d_counts + side, 1, ncclInt32, nbr, s->comm, st));
        NCCL_TRY(h, g_nccl.Recv(s->d_counts + 2 + side, 1, ncclInt32, nbr, s->comm, st));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    CU_TRY(h, cudaMemcpyAsync(s->h_ints, s->d_counts, 6 * sizeof(int), cudaMemcpyDeviceToHost, st));   // counts + the uniform flags
    CU_TRY(h, cudaStreamSynchronize(st));
    const int send_n[2] = {s->h_ints[0], s->h_ints[1]}, recv_n[2] = {s->h_ints[2], s->h_ints[3]};
    for (int side = 0; side < 2; side++)
        if (send_n[side] > s->buf_records || recv_n[side] > s->buf_records)
            return fail(h, SPH_E_CAPACITY, "slab exchange buffer too small");
    if ((long long)c.N + recv_n[0] + recv_n[1] > c.cap) return fail(h, SPH_E_CAPACITY, "slab particle capacity exceeded by migration / ghosts");
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int side = 0; side < 2; side++) {
        const int nbr = neighbour(s, side);
        if (nbr < 0) continue;
        if (send_n[side] > 0) NCCL_TRY(h, g_nccl.Send(s->sendbuf[side], (size_t)send_n[side] * REC_WORDS, ncclFloat32, nbr, s->comm, st));
        if (recv_n[side] > 0) NCCL_TRY(h, g_nccl.Recv(s->recvbuf[side], (size_t)recv_n[side] * REC_WORDS, ncclFloat32, nbr, s->comm, st));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    if (mode == 0 && c.N > 0) {   // leavers and the old ghosts die once the leavers are packed
        k_slab_retire<<<nb, SPH_BLOCK, 0, st>>>(c, h->d, s->z_lo, s->z_hi);
        h->launches++;
    }
    for (int side = 0; side < 2; side++) {
        if (recv_n[side] <= 0) continue;
        k_slab_unpack<<<(recv_n[side] + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, st>>>(h->d, s->recvbuf[side], recv_n[side], c.N, slot);
        c.N += recv_n[side];
        h->launches++;
    }
    s->halo_bytes += (int64_t)(send_n[0] + send_n[1]) * REC_WORDS * 4;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPH_OK : fail(h, SPH_E_CUDA, cudaGetErrorString(e));
}

}  // namespace

bool sph_is_slab(const SphHandle* h) { return h->slab != nullptr && h->slab->comm != nullptr; }

// Migration + ghost re-import, run by the sort path right before the cell histogram.
// Move slab boundaries by at most one cell layer each towards the busier rank.  Both ranks of a boundary decide from the
// same six numbers (their work, the work and size of the layer that would change hands, their free capacity, their
// thickness), so they always agree; the migration of the sort that follows carries the layer over.
static int slab_rebalance(SphHandle* h) {
    SlabState* s = h->slab;
    Consts& c = h->c;
    cudaStream_t st = h->stream;
    int* mine = s->d_work;
    CU_TRY(h, cudaMemsetAsync(mine, 0, 24 * sizeof(int), st));
    if (c.N > 0) {
        k_slab_work<<<(c.N + SPH_BLOCK - 1) / SPH_BLOCK, SPH_BLOCK, 0, st>>>(c, h->d, s->z_lo, s->z_hi, mine);
        h->launches++;
    }
    int* hv = s->h_ints + 28;   // [5] free capacity, [6] thickness: host-known, appended to the device figures
    hv[0] = c.cap - c.N;
    hv[1] = s->z_hi - s->z_lo;
    CU_TRY(h, cudaMemcpyAsync(mine + 5, hv, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int side = 0; side < 2; side++) {
        const int nbr = neighbour(s, side);
        if (nbr < 0) continue;
        NCCL_TRY(h, g_nccl.Send(mine, 8, ncclInt32, nbr, s->comm, st));
        NCCL_TRY(h, g_nccl.Recv(mine + 8 * (side + 1), 8, ncclInt32, nbr, s->comm, st));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    int v[24];
    CU_TRY(h, cudaMemcpyAsync(v, mine, sizeof v, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    // boundary between a lower rank A and an upper rank B: returns -1 (A gives its top layer to B), +1 (B gives its
    // bottom layer to A) or 0
    auto decide = [](const int* A, const int* B) {
        const long long diff = (long long)A[0] - B[0];
        if (diff > 0 && diff > A[2] && A[6] > 3 && B[5] > A[4] + 4096) return -1;
        if (diff < 0 && -diff > B[1] && B[6] > 3 && A[5] > B[3] + 4096) return +1;
        return 0;
    };
    int moved = 0;
    if (neighbour(s, 0) >= 0) {   // my lower boundary: the lower neighbour is A, I am B
        const int m = decide(v + 8, v);
        s->z_lo += m;
        moved += m != 0;
    }
    if (neighbour(s, 1) >= 0) {   // my upper boundary: I am A
        const int m = decide(v, v + 16);
        s->z_hi += m;
        moved += m != 0;
    }
    if (moved) {
        c.z_lo = s->z_lo;
        c.z_hi = s->z_hi;
        s->rebalances += moved;
    }
    return SPH_OK;
}

int sph_slab_pre_sort(SphHandle* h) {
    SlabState* s = h->slab;
    if (s->rebalance_period > 0 && s->world > 1 && h->rows_from_sort && (++s->sorts % s->rebalance_period) == 0) {
        const int rc0 = slab_rebalance(h);
        if (rc0) return rc0;
    }
    // rank-uniform flags (read back with the first count exchange): [4] dynamic rigid particles anywhere,
    // [5] boundary volumes stale anywhere
    s->h_ints[16] = 0;
    s->h_ints[17] = h->rigid_volume_clean ? 0 : 1;
    CU_TRY(h, cudaMemcpyAsync(s->d_counts + 4, s->h_ints + 16, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    sph_launch_count_dynamic_rigid(h, s->d_counts + 4);
    NCCL_TRY(h, g_nccl.AllReduce(s->d_counts + 4, s->d_counts + 4, 2, ncclInt32, ncclSum, s->comm, h->stream));
    int rc = exchange_phase(h, 0, SLOT_OWNED);
    if (rc) return rc;
    h->c.has_dynamic_rigid = s->h_ints[4] > 0 ? 1 : 0;
    h->dyn_rigid_dirty = false;
    if (s->h_ints[5] > 0) h->rigid_volume_clean = false;
    return exchange_phase(h, 1, SLOT_GHOST);
}

// After the scan: the live count and the layer boundaries (host needs them for launches / halos).
int sph_slab_post_scan(SphHandle* h) {
    SlabState* s = h->slab;
    Consts& c = h->c;
    const int plane = c.nx * c.ny;
    const int z[6] = {s->z_lo - 1, s->z_lo, s->z_lo + 1, s->z_hi - 1, s->z_hi, s->z_hi + 1};
    for (int k = 0; k < 6; k++) {
        const int zz = z[k] < 0 ? 0 : (z[k] > c.nz ? c.nz : z[k]);
        CU_TRY(h, cudaMemcpyAsync(s->h_ints + 8 + k, h->d.cell_start + (size_t)zz * plane, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CU_TRY(h, cudaMemcpyAsync(s->h_ints + 14, h->d.cell_start + c.ncell, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    const int* v = s->h_ints + 8;
    s->own_begin = v[1];
    s->own_end = v[4];
    s->ghost_lo_n = v[1] - v[0];
    s->ghost_hi_n = v[5] - v[4];
    s->send_lo_n = v[2] - v[1];
    s->send_hi_n = v[4] - v[3];
    if (neighbour(s, 0) < 0) { s->ghost_lo_n = 0; s->send_lo_n = 0; }
    if (neighbour(s, 1) < 0) { s->ghost_hi_n = 0; s->send_hi_n = 0; }
    c.N = s->h_ints[14];             // dead particles sit in the trash cell beyond the live ones
    c.row_begin = s->own_begin;
    c.row_end = s->own_end;
    h->rows_from_sort = true;
    sph_slab_publish_layout(h);
    return SPH_OK;
}

// Refresh one field of my ghosts from its owners: contiguous ranges, element = elem_bytes.
int sph_slab_halo(SphHandle* h, void* base, int elem_bytes) {
    if (!sph_is_slab(h)) return SPH_OK;
    SlabState* s = h->slab;
    char* b = (char*)base;
    const int lo = neighbour(s, 0), hi = neighbour(s, 1);
    NCCL_TRY(h, g_nccl.GroupStart());
    if (lo >= 0) {
        if (s->send_lo_n > 0) NCCL_TRY(h, g_nccl.Send(b + (size_t)s->own_begin * elem_bytes, (size_t)s->send_lo_n * elem_bytes, ncclInt8, lo, s->comm, h->stream));
        if (s->ghost_lo_n > 0) NCCL_TRY(h, g_nccl.Recv(b + (size_t)(s->own_begin - s->ghost_lo_n) * elem_bytes, (size_t)s->ghost_lo_n * elem_bytes, ncclInt8, lo, s->comm, h->stream));
    }
    if (hi >= 0) {
        if (s->send_hi_n > 0) NCCL_TRY(h, g_nccl.Send(b + (size_t)(s->own_end - s->send_hi_n) * elem_bytes, (size_t)s->send_hi_n * elem_bytes, ncclInt8, hi, s->comm, h->stream));
        if (s->ghost_hi_n > 0) NCCL_TRY(h, g_nccl.Recv(b + (size_t)s->own_end * elem_bytes, (size_t)s->ghost_hi_n * elem_bytes, ncclInt8, hi, s->comm, h->stream));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    s->halo_bytes += (int64_t)(s->send_lo_n + s->send_hi_n) * elem_bytes;
    s->halo_calls++;
    return SPH_OK;
}

// Sum a few doubles of Dev::red over all ranks (error sums of the solver loops).
int sph_slab_allreduce_red(SphHandle* h, int slot, int count) {
    if (!sph_is_slab(h)) return SPH_OK;
    NCCL_TRY(h, g_nccl.AllReduce(h->d.red + slot, h->d.red + slot, (size_t)count, ncclFloat64, ncclSum, h->slab->comm, h->stream));
    return SPH_OK;
}


// ---- peer memory ------------------------------------------------------------------------------------------
namespace {

struct PeerBlob {
    int magic, rank;
    cudaIpcMemHandle_t vm, vm_alt, aux, ctl;
};
static_assert(sizeof(PeerBlob) <= SPH_SLAB_PEER_BLOB_BYTES, "blob size");
constexpr int PEER_MAGIC = 0x53504831;

__device__ __forceinline__ int ld_volatile_i32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Ghost refresh straight from the neighbours' arrays: blockIdx.y = side (0: lower neighbour, 1: upper).  Waits until
// the neighbour has completed as many writers of this field as this rank has (its sweep's epilogue counts them), then
// copies the neighbour's boundary layer — same particles, same order as my ghost layer — over NVLink.
__global__ void __launch_bounds__(256) k_peer_pull(float4* mine_arr, const float4* peer_lo_arr, const float4* peer_hi_arr,
                                                   const PeerCtl* ctl, const PeerCtl* peer_lo, const PeerCtl* peer_hi, int which_aux,
                                                   int ghost_lo_begin, int ghost_lo_n, int ghost_hi_begin, int ghost_hi_n,
                                                   const double* red, int speculative) {
    if (speculative && red[CTRL_DONE] != 0.0) return;   // the writers returned early as well: nothing was signalled
    const int side = blockIdx.y;
    const PeerCtl* peer = side == 0 ? peer_lo : peer_hi;
    const int n = side == 0 ? ghost_lo_n : ghost_hi_n;
    if (!peer || n <= 0) return;
    if (threadIdx.x == 0) {
        const int want = which_aux ? ctl->aux_count : ctl->vel_count;
        const int* flag = which_aux ? &peer->aux_count : &peer->vel_count;
        while (ld_volatile_i32(flag) < want) __nanosleep(64);
        __threadfence_system();
    }
    __syncthreads();
    // my lower ghost layer is the lower neighbour's TOP boundary layer, my upper one the upper neighbour's BOTTOM layer
    const int src0 = side == 0 ? ld_volatile_i32(&peer->layout[1]) - n : ld_volatile_i32(&peer->layout[0]);
    const int dst0 = side == 0 ? ghost_lo_begin : ghost_hi_begin;
    const float4* src = side == 0 ? peer_lo_arr : peer_hi_arr;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) mine_arr[dst0 + k] = ld_volatile_f4(src + src0 + k);
}

}  // namespace

bool sph_slab_peers_ready(const SphHandle* h) { return h->slab && h->slab->peers_ready; }

PeerLinks sph_slab_peer_links(const SphHandle* h, int fuse_field) {
    PeerLinks l;
    memset(&l, 0, sizeof l);
    if (!(sph_slab_peers_ready(h) && h->peer_loop)) return l;
    const SlabState* s = h->slab;
    l.mine = s->ctl;
    l.all = s->d_peer_ctl;
    l.world = s->world;
    l.rank = s->rank;
    if (fuse_field) {
        const bool aux = fuse_field == GHOST_AUX;
        const int half = (h->d.vm == s->export_vm) ? 0 : 1;   // which half of the neighbours' velocity ping-pong is current
        const int nb[2] = {s->rank - 1, s->rank + 1};
        for (int side = 0; side < 2; side++) {
            if (nb[side] < 0 || nb[side] >= s->world) continue;
            const PeerCtl* pc = s->peer_ctl[nb[side]];
            l.ghost_arr[side] = aux ? s->peer_aux[side] : s->peer_vm[side][half];
            l.ghost_counter[side] = aux ? &pc->aux_count : &pc->vel_count;
            l.ghost_layout[side] = pc->layout;
        }
        l.my_counter = aux ? &s->ctl->aux_count : &s->ctl->vel_count;
        l.own_begin = s->own_begin;
        l.own_end = s->own_end;
        l.fused = 1;
    }
    return l;
}

// the index ranges the neighbours need to address my boundary layers (stream-ordered behind the sort)
void sph_slab_publish_layout(SphHandle* h) {
    SlabState* s = h->slab;
    if (!s || !s->ctl) return;
    int* v = s->h_ints + 24;
    v[0] = s->own_begin; v[1] = s->own_end; v[2] = s->send_lo_n; v[3] = s->send_hi_n;
    cudaMemcpyAsync(s->ctl->layout, v, 4 * sizeof(int), cudaMemcpyHostToDevice, h->stream);
}

void sph_slab_peer_pull(SphHandle* h, int which, bool speculative) {
    SlabState* s = h->slab;
    const bool aux = which == GHOST_AUX;
    const int lo = s->rank - 1, hi = s->rank + 1;
    // which half of the neighbours' velocity ping-pong is current: every rank has swapped as often as this one
    const int half = (h->d.vm == s->export_vm) ? 0 : 1;
    const float4* src_lo = lo >= 0 ? (aux ? s->peer_aux[0] : s->peer_vm[0][half]) : nullptr;
    const float4* src_hi = hi < s->world ? (aux ? s->peer_aux[1] : s->peer_vm[1][half]) : nullptr;
    const int nmax = s->ghost_lo_n > s->ghost_hi_n ? s->ghost_lo_n : s->ghost_hi_n;
    if (nmax <= 0) return;
    int gx = (nmax + 255) / 256;
    if (gx > 64) gx = 64;
    SphProf _prof(h, "k_peer_pull");
    k_peer_pull<<<dim3(gx, 2), 256, 0, h->stream>>>(aux ? h->d.aux : h->d.vm, src_lo, src_hi, s->ctl, lo >= 0 ? s->peer_ctl[lo] : nullptr,
                                                     hi < s->world ? s->peer_ctl[hi] : nullptr, aux ? 1 : 0, s->own_begin - s->ghost_lo_n,
                                                     s->ghost_lo_n, s->own_end, s->ghost_hi_n, h->d.red, speculative ? 1 : 0);
    h->launches++;
    s->peer_pulls++;
}

void sph_slab_free(SphHandle* h) {
    SlabState* s = h->slab;
    if (!s) return;
    for (int r = 0; r < SPH_PEER_MAX_RANKS; r++)
        if (s->peer_ctl[r] && s->peer_ctl[r] != s->ctl) cudaIpcCloseMemHandle(s->peer_ctl[r]);
    for (int side = 0; side < 2; side++) {
        for (int k = 0; k < 2; k++) if (s->peer_vm[side][k]) cudaIpcCloseMemHandle(s->peer_vm[side][k]);
        if (s->peer_aux[side]) cudaIpcCloseMemHandle(s->peer_aux[side]);
    }
    if (s->comm && g_nccl.ok) g_nccl.CommDestroy(s->comm);
    if (s->h_ints) cudaFreeHost(s->h_ints);
    delete s;
    h->slab = nullptr;
}

extern "C" {

int sph_slab_unique_id(void* out128) {
    if (!out128) return SPH_E_INVALID;
    if (!load_nccl()) return SPH_E_UNSUPPORTED;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return SPH_E_CUDA;
    memcpy(out128, &id, sizeof id);
    return SPH_OK;
}

int sph_slab_init(SphHandle* h, int32_t rank, int32_t world, const void* unique_id128, int32_t z_lo, int32_t z_hi,
                  int64_t global_particle_num) {
    if (!h || !unique_id128 || world < 1 || rank < 0 || rank >= world) return SPH_E_INVALID;
    if (!(h->P.flags & SPH_FLAG_SLAB)) return fail(h, SPH_E_STATE, "handle was not created with SPH_FLAG_SLAB");
    if (z_lo < 0 || z_hi > h->c.nz || z_lo >= z_hi) return fail(h, SPH_E_INVALID, "bad slab range");
    if (!load_nccl()) return fail(h, SPH_E_UNSUPPORTED, "libnccl.so.2 not found in the process");
    if (h->slab) sph_slab_free(h);
    SlabState* s = new SlabState();
    h->slab = s;
    s->rank = rank; s->world = world; s->z_lo = z_lo; s->z_hi = z_hi;
    h->c.z_lo = z_lo; h->c.z_hi = z_hi;
    h->c.ghost_lo = rank > 0 ? 1 : 0;
    h->c.ghost_hi = rank + 1 < world ? 1 : 0;
    cudaSetDevice(h->P.device);
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof id);
    NCCL_TRY(h, g_nccl.CommInitRank(&s->comm, world, id, rank));
    const size_t n = (size_t)h->c.cap;
    // a boundary layer holds at most a few percent of a slab; size the exchange for a quarter of it
    s->buf_records = (int)(n / 4 + 4096);
    void* p = nullptr;
#define DALLOC(ptr, bytes)                                     \
    CU_TRY(h, cudaMalloc(&p, (bytes)));                        \
    h->allocations.push_back(p);                               \
    ptr = (decltype(ptr))p
    DALLOC(s->flags, (n + 8) * sizeof(int));
    DALLOC(s->scan, (n + 8) * sizeof(int));
    for (int k = 0; k < 2; k++) {
        DALLOC(s->sendbuf[k], (size_t)s->buf_records * REC_WORDS * 4);
        DALLOC(s->recvbuf[k], (size_t)s->buf_records * REC_WORDS * 4);
    }
    DALLOC(s->d_counts, 8 * sizeof(int));
    DALLOC(s->d_work, 24 * sizeof(int));
    {
        const char* e = getenv("SPH_B200_REBALANCE");   // sorts between two re-balancing rounds, 0 = never
        s->rebalance_period = e ? atoi(e) : 64;
    }
#undef DALLOC
    CU_TRY(h, cudaMallocHost((void**)&s->h_ints, 32 * sizeof(int)));
    h->n_global = global_particle_num;
    h->c.row_begin = 0;
    h->c.row_end = h->c.N;
    h->sorted_valid = false;
    return SPH_OK;
}

int sph_slab_info(SphHandle* h, SphSlabInfo* out) {
    if (!h || !out) return SPH_E_INVALID;
    if (!h->slab) return fail(h, SPH_E_STATE, "not a slab handle");
    SlabState* s = h->slab;
    out->z_lo = s->z_lo; out->z_hi = s->z_hi;
    out->n_owned = s->own_end - s->own_begin;
    out->n_ghost = s->ghost_lo_n + s->ghost_hi_n;
    out->n_send_lo = s->send_lo_n; out->n_send_hi = s->send_hi_n;
    out->own_begin = s->own_begin; out->own_end = s->own_end;
    out->halo_bytes = s->halo_bytes; out->halo_calls = s->halo_calls;
    return SPH_OK;
}

int sph_slab_peer_export(SphHandle* h, void* blob) {
    if (!h || !blob) return SPH_E_INVALID;
    if (!h->slab) return fail(h, SPH_E_STATE, "not a slab handle");
    SlabState* s = h->slab;
    if (s->world > SPH_PEER_MAX_RANKS) return fail(h, SPH_E_UNSUPPORTED, "too many ranks for the peer control block");
    cudaSetDevice(h->P.device);
    if (!s->ctl) {
        void* p = nullptr;
        CU_TRY(h, cudaMalloc(&p, sizeof(PeerCtl)));
        h->allocations.push_back(p);
        s->ctl = (PeerCtl*)p;
        CU_TRY(h, cudaMemset(p, 0, sizeof(PeerCtl)));
        CU_TRY(h, cudaMalloc(&p, sizeof(PeerCtl*) * SPH_PEER_MAX_RANKS));
        h->allocations.push_back(p);
        s->d_peer_ctl = (PeerCtl**)p;
    }
    PeerBlob b;
    memset(&b, 0, sizeof b);
    b.magic = PEER_MAGIC;
    b.rank = s->rank;
    CU_TRY(h, cudaIpcGetMemHandle(&b.vm, h->d.vm));
    CU_TRY(h, cudaIpcGetMemHandle(&b.vm_alt, h->d.vm_alt));
    CU_TRY(h, cudaIpcGetMemHandle(&b.aux, h->d.aux));
    CU_TRY(h, cudaIpcGetMemHandle(&b.ctl, s->ctl));
    s->export_vm = h->d.vm;
    memset(blob, 0, SPH_SLAB_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return SPH_OK;
}

int sph_slab_peer_import(SphHandle* h, int32_t peer_rank, const void* blob) {
    if (!h || !blob) return SPH_E_INVALID;
    if (!h->slab || !h->slab->ctl) return fail(h, SPH_E_STATE, "export this rank's handles first");
    SlabState* s = h->slab;
    PeerBlob b;
    memcpy(&b, blob, sizeof b);
    if (b.magic != PEER_MAGIC || b.rank != peer_rank || peer_rank < 0 || peer_rank >= s->world || peer_rank == s->rank)
        return fail(h, SPH_E_INVALID, "bad peer blob");
    cudaSetDevice(h->P.device);
    void* p = nullptr;
    CU_TRY(h, cudaIpcOpenMemHandle(&p, b.ctl, cudaIpcMemLazyEnablePeerAccess));
    s->peer_ctl[peer_rank] = (PeerCtl*)p;
    const int side = peer_rank == s->rank - 1 ? 0 : (peer_rank == s->rank + 1 ? 1 : -1);
    if (side >= 0) {
        CU_TRY(h, cudaIpcOpenMemHandle(&p, b.vm, cudaIpcMemLazyEnablePeerAccess));
        s->peer_vm[side][0] = (float4*)p;
        CU_TRY(h, cudaIpcOpenMemHandle(&p, b.vm_alt, cudaIpcMemLazyEnablePeerAccess));
        s->peer_vm[side][1] = (float4*)p;
        CU_TRY(h, cudaIpcOpenMemHandle(&p, b.aux, cudaIpcMemLazyEnablePeerAccess));
        s->peer_aux[side] = (float4*)p;
    }
    if (++s->imported == s->world - 1) {
        s->peer_ctl[s->rank] = s->ctl;
        CU_TRY(h, cudaMemcpy(s->d_peer_ctl, s->peer_ctl, sizeof(PeerCtl*) * SPH_PEER_MAX_RANKS, cudaMemcpyHostToDevice));
        s->peers_ready = true;
    }
    return SPH_OK;
}

int sph_slab_set_global_particle_num(SphHandle* h, int64_t n) {
    if (!h) return SPH_E_INVALID;
    h->n_global = n;
    return SPH_OK;
}

}  // extern "C"
