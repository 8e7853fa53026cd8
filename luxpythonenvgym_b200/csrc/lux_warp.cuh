// lux_warp.cuh -- lane-group primitives used by the turn engine.
//
// A match is played by a GROUP of G consecutive lanes of a warp (G = 8, 16 or 32, a template
// parameter; 32 / G matches share a warp).  A typical match holds a handful of units and city
// tiles, so a full warp per match leaves most lanes idle in most phases; narrow groups put
// several matches behind every issued instruction.
//
// The groups of a warp run in LOCKSTEP: every collective below is executed by all 32 lanes with
// the full member mask (a per-group member mask would make the hardware run the groups one after
// the other), and the group's own result is cut out of the warp-wide one.  The engine therefore
// keeps every collective in warp-uniform control flow: loops that contain one run for the
// largest trip count of the warp (lux_max_warp) and phase-level conditions are decided with
// lux_any_warp; lanes of a group that has nothing to do are simply predicated off.
// Lane numbers and ballot bits are group-relative.
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
#define LUX_FULL 0xffffffffu

LUX_DEV int lux_lane32() { return (int)(threadIdx.x & 31u); }
template <int G = 32> LUX_DEV unsigned lux_group_shift() { return G == 32 ? 0u : (threadIdx.x & 31u) & ~(unsigned)(G - 1); }
template <int G = 32> LUX_DEV int lux_lane() { return (int)(threadIdx.x & (unsigned)(G - 1)); }
LUX_DEV void lux_sync_warp() { __syncwarp(); }
template <int G = 32> LUX_DEV unsigned lux_ballot(int pred) {
    unsigned m = __ballot_sync(LUX_FULL, pred);
    return G == 32 ? m : (m >> lux_group_shift<G>()) & ((1u << (G & 31)) - 1u);
}
LUX_DEV int lux_any_warp(int pred) { return __any_sync(LUX_FULL, pred); }
LUX_DEV int lux_max_warp(int v) { return __reduce_max_sync(LUX_FULL, v); }
template <int G = 32> LUX_DEV int lux_shfl(int v, int src) { return __shfl_sync(LUX_FULL, v, src, G); }
template <int G = 32> LUX_DEV int lux_reduce_add(int v) {
    if (G == 32) return __reduce_add_sync(LUX_FULL, v);
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(LUX_FULL, v, d, G);
    return v;
}
// lanes of the caller's group holding the same key (keys must stay below 2^26)
template <int G = 32> LUX_DEV unsigned lux_match_any(int key) {
    unsigned m = __match_any_sync(LUX_FULL, G == 32 ? key : (int)((unsigned)key | (lux_group_shift<G>() << 26)));
    return G == 32 ? m : (m >> lux_group_shift<G>()) & ((1u << (G & 31)) - 1u);
}
LUX_DEV unsigned lux_reduce_min_u(unsigned v) { return __reduce_min_sync(LUX_FULL, v); }     // whole warp
LUX_DEV unsigned lux_reduce_max_u(unsigned v) { return __reduce_max_sync(LUX_FULL, v); }
LUX_DEV int lux_popc(unsigned v) { return __popc(v); }
LUX_DEV int lux_ffs(unsigned v) { return __ffs(v); }
LUX_DEV unsigned lux_atomic_or(unsigned* p, unsigned v) { return atomicOr(p, v); }
LUX_DEV void lux_atomic_and(unsigned* p, unsigned v) { atomicAnd(p, v); }
LUX_DEV int lux_atomic_add(int* p, int v) { return atomicAdd(p, v); }
#endif

// warp-wide decisions over GROUP-UNIFORM arguments (every lane of a group passes the same value): with one
// group per warp they are the argument itself
template <int G = 32> LUX_DEV int lux_wany(int group_pred) { return G == 32 ? group_pred : lux_any_warp(group_pred); }
template <int G = 32> LUX_DEV int lux_wmax(int group_v) { return G == 32 ? group_v : lux_max_warp(group_v); }

// group-level shorthands (all of them warp-wide collectives underneath)
template <int G = 32> LUX_DEV void lux_sync() { lux_sync_warp(); }
template <int G = 32> LUX_DEV int lux_any(int pred) { return lux_ballot<G>(pred) != 0; }
template <int G = 32> LUX_DEV unsigned lux_lanemask_lt() { return (1u << lux_lane<G>()) - 1u; }

// exclusive scan of small ints over one G-wide chunk; *total = the chunk's sum
template <int G = 32> LUX_DEV int lux_exscan_add(int v, int* total) {
    int lane = lux_lane<G>();
    int x = v;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        int y = lux_shfl<G>(x, lane - d < 0 ? lane : lane - d);
        if (lane >= d) x += y;
    }
    *total = lux_shfl<G>(x, G - 1);
    return x - v;
}

// byte-granular atomic OR inside a 4-byte aligned byte plane; returns the old byte
LUX_DEV unsigned lux_atomic_or_byte(uint8_t* plane, int idx, unsigned bits) {
    unsigned* w = (unsigned*)(plane + (idx & ~3));
    int sh = (idx & 3) * 8;
    unsigned old = lux_atomic_or(w, bits << sh);
    return (old >> sh) & 0xffu;
}
