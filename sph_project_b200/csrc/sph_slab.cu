// sph_slab.cu — Z-slab sharding entry points; filled in by the multi-GPU milestone.
#include "sph_kernels.h"

static int unsupported(SphHandle* h) {
    if (h) h->err = "Z-slab sharding is not built yet";
    return SPH_E_UNSUPPORTED;
}

extern "C" {
int sph_slab_set_range(SphHandle* h, int32_t, int32_t) { return unsupported(h); }
int sph_slab_info(SphHandle* h, SphSlabInfo*) { return unsupported(h); }
int sph_slab_begin_exchange(SphHandle* h, int32_t*) { return unsupported(h); }
int sph_slab_pack(SphHandle* h, int32_t, int32_t, void**, int32_t*) { return unsupported(h); }
int sph_slab_unpack(SphHandle* h, int32_t, int32_t, const void*, int32_t) { return unsupported(h); }
int sph_slab_halo_pack(SphHandle* h, int32_t, int32_t, void**, int32_t*, int32_t*) { return unsupported(h); }
int sph_slab_halo_unpack(SphHandle* h, int32_t, int32_t, const void*, int32_t) { return unsupported(h); }
int sph_slab_halo_recv_count(SphHandle* h, int32_t, int32_t*) { return unsupported(h); }
}
