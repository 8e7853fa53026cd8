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
};

inline State& st() {
    static thread_local State s;
    return s;
}

inline void yield_next() {
    State& s = st();
    int from = s.cur;
    for (int k = 1; k <= 32; k++) {
        int nxt = (from + k) & 31;
        if (!s.done[nxt]) {
            if (nxt == from) return;
            s.cur = nxt;
            swapcontext(&s.ctx[from], &s.ctx[nxt]);
            return;
        }
    }
}

inline void barrier() {
    State& s = st();
    unsigned gen = s.gen;
    if (++s.arrived == 32) {
        s.arrived = 0;
        s.gen++;
        return;
    }
    while (s.gen == gen) yield_next();
}

inline void trampoline() {
    State& s = st();
    s.fn(s.arg);
    s.done[s.cur] = 1;
    // hand over to another live lane, or back to main when all are finished
    for (int k = 1; k <= 32; k++) {
        int nxt = (s.cur + k) & 31;
        if (!s.done[nxt]) {
            int from = s.cur;
            s.cur = nxt;
            swapcontext(&s.ctx[from], &s.ctx[nxt]);
        }
    }
    setcontext(&s.main_ctx);
}

// Runs fn(arg) on 32 lanes; returns when every lane has returned.
inline void run_warp(void (*fn)(void*), void* arg) {
    State& s = st();
    const size_t STACK = 256 * 1024;
    s.fn = fn;
    s.arg = arg;
    s.arrived = 0;
    for (int i = 0; i < 32; i++) {
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

static inline int lux_lane() { return warp_emu::st().cur; }
static inline void lux_sync() { warp_emu::barrier(); }
static inline unsigned lux_ballot(int pred) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = pred ? 1 : 0;
    warp_emu::barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(s.vals[i] != 0) << i;
    warp_emu::barrier();
    return m;
}
static inline int lux_any(int pred) { return lux_ballot(pred) != 0; }
static inline int lux_shfl(int v, int src) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = v;
    warp_emu::barrier();
    int r = (int)s.vals[src & 31];
    warp_emu::barrier();
    return r;
}
static inline int lux_reduce_add(int v) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = v;
    warp_emu::barrier();
    long long r = 0;
    for (int i = 0; i < 32; i++) r += s.vals[i];
    warp_emu::barrier();
    return (int)r;
}
static inline unsigned lux_match_any(int key) {
    warp_emu::State& s = warp_emu::st();
    s.vals[s.cur] = key;
    warp_emu::barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(s.vals[i] == key) << i;
    warp_emu::barrier();
    return m;
}
static inline int lux_popc(unsigned v) { return __builtin_popcount(v); }
static inline int lux_ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned lux_atomic_or(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int lux_atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
