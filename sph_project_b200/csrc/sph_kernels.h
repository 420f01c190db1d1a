// sph_kernels.h — host-side launchers of every kernel + the block reduction helper.
#pragma once

#include "sph_common.cuh"

// slots of Dev::red (double), see sph_stream.cu
enum RedSlot {
    RED_ERR = 0,        // sum for DFSPH / PCISPH error reductions
    RED_MASS = 1,       // compute_rigid_body_mass
    RED_CG_RR = 2,      // sum |r|^2 (numerator of alpha)
    RED_CG_PAP = 3,     // sum p . Ap
    RED_CG_RR_NEW = 4,  // sum |r_new|^2 (numerator of beta, cg_error^2)
    RED_CG_RR_OLD = 5,  // sum |r_old|^2 (denominator of beta)
    // loop control of the DFSPH solves when the exit test runs on the device (sph_sweeps.cu, SolveMode)
    CTRL_DONE = 32,     // != 0: the solve has converged; speculative launches return at once
    CTRL_ITERS = 33,    // iterations tested so far
    CTRL_ERR = 34,      // average error of the last tested iteration
    CTRL_COUNT = 3,
    RED_COUNT = 48
};


#ifdef __CUDACC__
// sum `v` over the block, one atomicAdd(double) per block
__device__ __forceinline__ void block_reduce_add(double* dst, double v) {
    __shared__ double warp_part[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double s = lane < nw ? warp_part[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s != 0.0) atomicAdd(dst, s);
    }
    __syncthreads();
}
#endif

// ---- peer memory of the Z-slab solver loops (sph_slab.cu) -------------------------------------------------
// One control block per rank in device memory, mapped into every other rank through CUDA IPC.  Counters only ever
// grow; all ranks run the same kernels in the same order, so "the peer has signalled as often as I have" is the
// whole handshake.
#define SPH_PEER_MAX_RANKS 16
struct PeerCtl {
    int vel_count;                       // velocity writers (correction sweeps) this rank has completed
    int aux_count;                       // payload writers (density-change sweeps)
    int err_count;                       // error sums this rank has delivered to everybody
    int pad0;
    int layout[4];                       // own_begin, own_end, send_lo_n, send_hi_n after the last sort
    int err_flag[SPH_PEER_MAX_RANKS];    // written by rank q: error sums q has delivered into THIS block
    double partial[2][SPH_PEER_MAX_RANKS];   // [delivery parity][rank q]: q's error sum
};
// what a sweep needs to signal / deliver in its epilogue (all null / 0 outside peer loops)
struct PeerLinks {
    PeerCtl* mine;
    PeerCtl* const* all;                 // device array [world] of every rank's block (own included)
    int world, rank;
    // fused ghost reads: the sweep stages the ghost runs of its payload array straight from the neighbours' live arrays
    int fused;                           // 0: ghosts were refreshed into the local array before the launch
    const float4* ghost_arr[2];          // lower / upper neighbour's payload array (null: no neighbour)
    const int* ghost_counter[2];         // their completion counters for that payload
    const int* ghost_layout[2];          // their layout words (own_begin, own_end, ...)
    const int* my_counter;               // this rank's counter: the neighbours must have counted as far
    int own_begin, own_end;              // my owned index range: ghosts lie below / above it
};
bool sph_slab_peers_ready(const SphHandle* h);
PeerLinks sph_slab_peer_links(const SphHandle* h, int fuse_field = 0 /* GHOST_VEL | GHOST_AUX: read that payload's ghosts from the neighbours */);
void sph_slab_peer_pull(SphHandle* h, int which /* GHOST_VEL | GHOST_AUX */, bool speculative);
void sph_slab_publish_layout(SphHandle* h);

// sph_slab.cu
enum GhostField { GHOST_VEL = 1, GHOST_AUX = 2, GHOST_RHO = 4, GHOST_PV = 8 };
bool sph_is_slab(const SphHandle* h);
int sph_slab_pre_sort(SphHandle* h);
int sph_slab_post_scan(SphHandle* h);
int sph_slab_halo(SphHandle* h, void* base, int elem_bytes);
int sph_slab_allreduce_red(SphHandle* h, int slot, int count);
void sph_slab_free(SphHandle* h);
// mark / refresh ghost copies (no-ops unless the handle is a slab)
inline void sph_ghost_dirty(SphHandle* h, int what) {
    if (!h->slab) return;
    h->ghost_stale |= what;
    h->peer_signalled &= ~what;   // a writer that does not signal: the next refresh goes through NCCL (peer-loop launchers set the bit afterwards)
}
void sph_ghost_sync(SphHandle* h, int what);

// sph_grid.cu
void sph_exclusive_scan(SphHandle* h, const int* in, int n, int* out);
int sph_sort_particles(SphHandle* h);
void sph_bricks_refresh(SphHandle* h);   // brick work list after host edits of materials

// sph_sweeps.cu
void sph_launch_rigid_volume(SphHandle* h);
void sph_launch_density(SphHandle* h, bool with_alpha = false);   // with_alpha: DFSPH compute_alpha on the same staged window
void sph_launch_pressure_accel(SphHandle* h);
void sph_launch_temp_pressure_accel(SphHandle* h);
void sph_launch_surface_tension(SphHandle* h);
void sph_launch_viscosity(SphHandle* h);
void sph_launch_dfsph_alpha(SphHandle* h);
// fused: + kappa(_v) + aux payload; mode (SolveMode: 0 plain, 1 error sum, 2 error sum + exit test on the device),
// speculative: return at once when an earlier launch of the batch has converged
void sph_launch_dfsph_density_derivative(SphHandle* h, bool fused, int mode = 0, bool speculative = false, float eta = 0.0f);
void sph_launch_dfsph_density_star(SphHandle* h, bool fused, int mode = 0, bool speculative = false, float eta = 0.0f);
bool sph_lists_ready(SphHandle* h);
void sph_launch_dfsph_solve_check(SphHandle* h, float eta);   // Z-slabs: the exit test after the all-reduce of the error sum
void sph_launch_dfsph_correct_divergence(SphHandle* h, bool aux_ready, bool speculative = false);   // aux_ready: fused kernel wrote kappa into aux
void sph_launch_dfsph_correct_density(SphHandle* h, bool aux_ready, bool speculative = false);
void sph_launch_pcisph_density_star(SphHandle* h);
void sph_launch_cg_prepare1(SphHandle* h);
void sph_launch_cg_Ap(SphHandle* h, bool aux_ready);
void sph_launch_neighbor_count(SphHandle* h, int* counts);
void sph_launch_neighbor_fill(SphHandle* h, const int* offsets, int* indices);

// sph_stream.cu
void sph_launch_gravity(SphHandle* h);
void sph_launch_update_velocity(SphHandle* h);
void sph_launch_update_position(SphHandle* h);
void sph_launch_boundary(SphHandle* h, int particle_type);
void sph_launch_renew_rigid(SphHandle* h);
void sph_launch_prepare_emitter(SphHandle* h);
void sph_launch_wcsph_pressure(SphHandle* h);
void sph_launch_dfsph_kappa_v(SphHandle* h);
void sph_launch_dfsph_kappa(SphHandle* h);
void sph_launch_dfsph_divergence_error(SphHandle* h);   // red[RED_ERR] = sum rho0 * drho (fluid)
void sph_launch_dfsph_density_error(SphHandle* h);      // red[RED_ERR] = sum rho_star - 1 (fluid)
void sph_launch_pcisph_predict_velocity(SphHandle* h);
void sph_launch_pcisph_predict_position(SphHandle* h);
void sph_launch_pcisph_update_pressure(SphHandle* h);
void sph_launch_pcisph_init_step(SphHandle* h);
void sph_launch_cg_prepare1_pre(SphHandle* h);
void sph_launch_cg_prepare2(SphHandle* h);
void sph_launch_cg_dots(SphHandle* h);
void sph_launch_cg_update_x(SphHandle* h);
void sph_launch_cg_update_r(SphHandle* h);
void sph_launch_cg_update_p(SphHandle* h);
void sph_launch_cg_prepare_guess(SphHandle* h);
void sph_launch_cg_velocity_from_x(SphHandle* h);
void sph_launch_cg_velocity_restore(SphHandle* h);
void sph_launch_rigid_body_mass(SphHandle* h, int object_id);
void sph_launch_count_dynamic_rigid(SphHandle* h, int* out_dev);
void sph_fill_i32(SphHandle* h, int* p, size_t n, int v);
void sph_fill_f32(SphHandle* h, float* p, size_t n, float v);
// field <-> dense staging conversion (host layout: [n, comps] f32 / i32)
int sph_field_to_staging(SphHandle* h, int field, int n);
int sph_staging_to_field(SphHandle* h, int field, int n);
void sph_launch_cell_coords(SphHandle* h, int* out, int n);
void sph_launch_ref_cell_hist(SphHandle* h, int* hist);
