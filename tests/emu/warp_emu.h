// warp_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// Host emulation of one CUDA warp: 32 cooperative fibers (ucontext), switched round-robin
// whenever a lane reaches a warp collective.  A lane therefore runs as far ahead of the
// others as the CUDA execution model allows (up to the next __syncwarp / ballot / shuffle),
// which is the schedule most likely to expose a missing synchronisation.  Collectives are
// always full-mask, exactly as the kernel uses them.  This exists so the CUDA engine's
// source (luxpythonenvgym_b200/csrc/lux_step_core.cuh) can be compiled with g++ and
// compared with the oracle in a container that has no GPU.  It is never part of the product.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

namespace warp_emu {

struct State {
    ucontext_t main_ctx;
    ucontext_t ctx[32];
    char* stacks[32];
    int done[32];
    int cur;
    int arrived;
    unsigned gen;
    long long vals[32];
    void (*fn)(void*);
    void* arg;
    int kind;       // collective the lanes of the current barrier generation are in
    int width;      // live fibers (the lane-group width G)
};

inline State& st() {
    static thread_local State s;
    return s;
}

inline void yield_next() {
    State& s = st();
    int from = s.cur;
    for (int k = 1; k <= s.width; k++) {
        int nxt = (from + k) % s.width;
        if (!s.done[nxt]) {
            if (nxt == from) return;
            s.cur = nxt;
            swapcontext(&s.ctx[from], &s.ctx[nxt]);
            return;
        }
    }
}

inline void barrier(int kind = 0) {
    State& s = st();
    unsigned gen = s.gen;
    // every lane of the warp must be at the same kind of collective (sync / ballot / shuffle / ...)
    if (s.arrived == 0) s.kind = kind;
    else if (s.kind != kind) {
        fprintf(stderr, "warp_emu: lanes meet at different collectives (%d vs %d): non-uniform control flow\n", s.kind, kind);
        abort();
    }
    if (++s.arrived == s.width) {
        s.arrived = 0;
        s.gen++;
        return;
    }
    while (s.gen == gen) {
        int live = 0;
        for (int i = 0; i < s.width; i++) live += !s.done[i];
        if (live < s.width) {   // a lane left the kernel while others wait at a collective
            fprintf(stderr, "warp_emu: collective reached by %d of %d lanes only (non-uniform control flow)\n", live, s.width);
            abort();
        }
        yield_next();
    }
}

inline void trampoline() {
    State& s = st();
    s.fn(s.arg);
    s.done[s.cur] = 1;
    // hand over to another live lane, or back to main when all are finished
    for (int k = 1; k <= s.width; k++) {
        int nxt = (s.cur + k) % s.width;
        if (!s.done[nxt]) {
            int from = s.cur;
            s.cur = nxt;
            swapcontext(&s.ctx[from], &s.ctx[nxt]);
        }
    }
    setcontext(&s.main_ctx);
}

// Runs fn(arg) on `width` lanes (one lane group); returns when every lane has returned.
inline void run_warp(void (*fn)(void*), void* arg, int width = 32) {
    State& s = st();
    const size_t STACK = 256 * 1024;
    s.fn = fn;
    s.arg = arg;
    s.arrived = 0;
    s.width = width;
    for (int i = 0; i < width; i++) {
        if (!s.stacks[i]) s.stacks[i] = (char*)malloc(STACK);
        s.done[i] = 0;
        getcontext(&s.ctx[i]);
        s.ctx[i].uc_stack.ss_sp = s.stacks[i];
        s.ctx[i].uc_stack.ss_size = STACK;
        s.ctx[i].uc_link = &s.main_ctx;
        makecontext(&s.ctx[i], (void (*)())trampoline, 0);
    }
    s.cur = 0;
    swapcontext(&s.main_ctx, &s.ctx[0]);
}

}  // namespace warp_emu

// ---- the lane-group primitives of lux_warp.cuh: 32 fibers = one warp of 32 / G lane groups ------
static inline int lux_lane32() { return warp_emu::st().cur; }
template <int G = 32> static inline int lux_lane() { return warp_emu::st().cur & (G - 1); }
static inline void lux_sync_warp() { warp_emu::barrier(1); }
template <int G = 32> static inline unsigned lux_ballot(int pred) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = pred ? 1 : 0;
    warp_emu::barrier(2);
    unsigned m = 0;
    int base = s.cur & ~(G - 1);
    for (int i = 0; i < G; i++) m |= (unsigned)(s.vals[base + i] != 0) << i;
    warp_emu::barrier(2);
    return m;
}
static inline int lux_any_warp(int pred) { return lux_ballot<32>(pred) != 0; }
static inline int lux_max_warp(int v) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = v;
    warp_emu::barrier(3);
    long long r = s.vals[0];
    for (int i = 1; i < 32; i++) if (s.vals[i] > r) r = s.vals[i];
    warp_emu::barrier(3);
    return (int)r;
}
template <int G = 32> static inline int lux_shfl(int v, int src) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = v;
    warp_emu::barrier(4);
    int r = (int)s.vals[(s.cur & ~(G - 1)) + (src & (G - 1))];
    warp_emu::barrier(4);
    return r;
}
template <int G = 32> static inline int lux_reduce_add(int v) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = v;
    warp_emu::barrier(5);
    long long r = 0;
    int base = s.cur & ~(G - 1);
    for (int i = 0; i < G; i++) r += s.vals[base + i];
    warp_emu::barrier(5);
    return (int)r;
}
template <int G = 32> static inline unsigned lux_match_any(int key) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = key;
    warp_emu::barrier(6);
    unsigned m = 0;
    int base = s.cur & ~(G - 1);
    for (int i = 0; i < G; i++) m |= (unsigned)(s.vals[base + i] == key) << i;
    warp_emu::barrier(6);
    return m;
}
static inline unsigned lux_reduce_min_u(unsigned v) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = (long long)v;
    warp_emu::barrier(7);
    long long r = s.vals[0];
    for (int i = 1; i < 32; i++) if (s.vals[i] < r) r = s.vals[i];
    warp_emu::barrier(7);
    return (unsigned)r;
}
static inline unsigned lux_reduce_max_u(unsigned v) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = (long long)v;
    warp_emu::barrier(8);
    long long r = s.vals[0];
    for (int i = 1; i < 32; i++) if (s.vals[i] > r) r = s.vals[i];
    warp_emu::barrier(8);
    return (unsigned)r;
}
static inline int lux_popc(unsigned v) { return __builtin_popcount(v); }
static inline int lux_ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned lux_atomic_or(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline void lux_atomic_and(unsigned* p, unsigned v) { *p &= v; }
static inline int lux_atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
