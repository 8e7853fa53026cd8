// lux_warp.cuh -- warp-level primitives used by the turn engine.
//
// On the device these are the sm_100a intrinsics.  When LUX_EMULATE is defined (only by
// tests/emu/, never by the product build) the same kernel source is compiled for the host
// and the 32 lanes of a warp run as cooperative fibers (tests/emu/warp_emu.h) so the
// engine's logic can be checked against the oracle in a container without a GPU.
#pragma once
#include <stdint.h>

#ifdef LUX_EMULATE
#include "warp_emu.h"
#define LUX_DEV static inline
#else
#define LUX_DEV __device__ __forceinline__

LUX_DEV int lux_lane() { return (int)(threadIdx.x & 31); }
LUX_DEV void lux_sync() { __syncwarp(); }
LUX_DEV unsigned lux_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
LUX_DEV int lux_any(int pred) { return __any_sync(0xffffffffu, pred); }
LUX_DEV int lux_shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
LUX_DEV int lux_reduce_add(int v) { return __reduce_add_sync(0xffffffffu, v); }
LUX_DEV int lux_popc(unsigned v) { return __popc(v); }
LUX_DEV unsigned lux_match_any(int key) { return __match_any_sync(0xffffffffu, key); }
LUX_DEV int lux_ffs(unsigned v) { return __ffs(v); }
LUX_DEV unsigned lux_atomic_or(unsigned* p, unsigned v) { return atomicOr(p, v); }
LUX_DEV int lux_atomic_add(int* p, int v) { return atomicAdd(p, v); }
#endif

LUX_DEV unsigned lux_lanemask_lt() { return (1u << lux_lane()) - 1u; }

// inclusive-exclusive scan of small ints over one 32-wide chunk
LUX_DEV int lux_exscan_add(int v, int* total) {
    int lane = lux_lane();
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = lux_shfl(x, lane - d < 0 ? lane : lane - d);
        if (lane >= d) x += y;
    }
    *total = lux_shfl(x, 31);
    return x - v;
}

// byte-granular atomic OR inside a 4-byte aligned byte plane; returns the old byte
LUX_DEV unsigned lux_atomic_or_byte(uint8_t* plane, int idx, unsigned bits) {
    unsigned* w = (unsigned*)(plane + (idx & ~3));
    int sh = (idx & 3) * 8;
    unsigned old = lux_atomic_or(w, bits << sh);
    return (old >> sh) & 0xffu;
}
