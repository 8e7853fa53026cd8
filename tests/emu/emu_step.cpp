// emu_step.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the CUDA turn engine's source (luxpythonenvgym_b200/csrc/lux_step_core.cuh) for the
// host with the warp emulator of warp_emu.h and exposes it with the same batch signature as
// lux_b200_step, so tests can diff the kernel's logic against the oracle without a GPU.
// Built by tests/emu/build.py into tests/emu/_build/ ; never loaded by the product.
#define LUX_EMULATE 1
#include <stdlib.h>
#include <string.h>

#include "../../luxpythonenvgym_b200/csrc/lux_step_core.cuh"

struct EmuArgs {
    const LuxBatch* b;
    int m;
    const uint32_t* ua;
    const uint8_t* ta;
    unsigned char* smem;
    LuxSmemLayout L;
    unsigned flags;
    int dirty;
    int group;
};

template <int G>
static void emu_lane(void* p) {
    EmuArgs* a = (EmuArgs*)p;
    const LuxBatch* b = a->b;
    const int grp = lux_lane32() / G;
    int m = a->m + grp;                       // 32 / G matches per warp, one lane group each
    int valid = m < b->n_matches;
    if (!valid) m = 0;
    int ncell = b->size * b->size;
    LuxGlobalMatch g;
    g.hdr = b->hdr + m;
    g.cells = b->cells + (size_t)m * ncell;
    g.units = b->units + (size_t)m * b->u_cap;
    g.cities = b->cities + (size_t)m * b->c_cap;
    g.tiles = b->tiles + (size_t)m * b->t_cap;
    g.res = b->res + (size_t)m * b->r_cap;
    g.uact = a->ua + (size_t)m * b->u_cap;
    g.tact = a->ta + (size_t)m * b->t_cap;
    unsigned char* smem = a->smem + (size_t)grp * a->L.total;
    lux_zero16<G>(smem + a->L.p1, (ncell + 15) & ~15);
    lux_zero16<G>(smem + a->L.occ, (ncell + 15) & ~15);
    LuxMatch M;
    int result;
    int active = lux_stage_plain<G>(M, smem, g, valid, b->size, b->u_cap, b->c_cap, b->t_cap, b->r_cap, a->L.total, &result);
    LuxTurnOut T = b->size == 32 ? lux_turn<32, G>(M, a->flags) : lux_turn<0, G>(M, a->flags);
    lux_finish_and_store<0, G>(M, T, g.hdr, g.units, g.cities, g.tiles, g.res, active);
    lux_sync<G>();
    // the engine's contract: both byte planes are clean again after the turn
    for (int i = lux_lane<G>(); i < ncell; i += G)
        if (M.p1[i] || M.occ[i]) a->dirty = 1;
}

extern "C" size_t lux_emu_smem_bytes(int size, int u_cap, int c_cap, int t_cap, int r_cap) {
    return (size_t)lux_smem_layout(size, u_cap, c_cap, t_cap, r_cap).total;
}

extern "C" int lux_emu_step_group(const LuxBatch* b, const uint32_t* unit_actions, const uint8_t* tile_actions, unsigned flags,
                                  int group) {
    if (!b || !unit_actions || !tile_actions) return LUX_E_ARG;
    if (group != 8 && group != 16 && group != 32) return LUX_E_ARG;
    EmuArgs a;
    a.b = b; a.ua = unit_actions; a.ta = tile_actions; a.flags = flags; a.dirty = 0; a.group = group;
    a.L = lux_smem_layout(b->size, b->u_cap, b->c_cap, b->t_cap, b->r_cap);
    const int gpw = 32 / group;
    a.smem = (unsigned char*)aligned_alloc(16, (size_t)a.L.total * gpw);
    for (int m = 0; m < b->n_matches; m += gpw) {
        memset(a.smem, 0xA5, (size_t)a.L.total * gpw);   // shared memory is not zero-initialised on the device either
        a.m = m;
        warp_emu::run_warp(group == 8 ? emu_lane<8> : group == 16 ? emu_lane<16> : emu_lane<32>, &a, 32);
    }
    free(a.smem);
    return a.dirty ? -100 : 0;
}

// the per-turn shared-memory need of a header (lux_need_header + lux_smem_layout_need), for property tests
extern "C" int lux_emu_need(const LuxHeader* h, int size, int u_cap, int c_cap, int t_cap, int r_cap, int* out5) {
    LuxNeed n = lux_need_header(h, u_cap, c_cap, t_cap, r_cap);
    out5[0] = n.u; out5[1] = n.c; out5[2] = n.t; out5[3] = n.r; out5[4] = n.q;
    return lux_smem_layout_need(size, n).total;
}

// one turn with `slice` bytes per match: matches whose turn does not fit are left untouched (returns how many)
extern "C" int lux_emu_step_slice(const LuxBatch* b, const uint32_t* unit_actions, const uint8_t* tile_actions, unsigned flags,
                                  int slice) {
    if (!b || !unit_actions || !tile_actions) return LUX_E_ARG;
    int skipped = 0;
    unsigned char* smem = (unsigned char*)aligned_alloc(16, (size_t)slice + 64);
    for (int m = 0; m < b->n_matches; m++) {
        const LuxHeader* h = b->hdr + m;
        if (!h->done && lux_smem_layout_need(b->size, lux_need_header(h, b->u_cap, b->c_cap, b->t_cap, b->r_cap)).total > slice) skipped++;
    }
    EmuArgs a;
    a.b = b; a.ua = unit_actions; a.ta = tile_actions; a.flags = flags; a.dirty = 0; a.group = 32;
    a.L = lux_smem_layout(b->size, b->u_cap, b->c_cap, b->t_cap, b->r_cap);
    a.L.total = slice;                         // emu_lane sizes and limits the slice by L.total
    a.smem = smem;
    for (int m = 0; m < b->n_matches; m++) {
        memset(smem, 0xA5, (size_t)slice);
        memset(smem + slice, 0x5A, 64);        // canary behind the slice
        a.m = m;
        warp_emu::run_warp(emu_lane<32>, &a, 32);
        for (int i = 0; i < 64; i++)
            if (smem[slice + i] != 0x5A) { free(smem); return -101; }
    }
    free(smem);
    return a.dirty ? -100 : skipped;
}

extern "C" int lux_emu_step(const LuxBatch* b, const uint32_t* unit_actions, const uint8_t* tile_actions, unsigned flags) {
    return lux_emu_step_group(b, unit_actions, tile_actions, flags, 32);
}

// ---- observations through the emulated warp ------------------------------------------------
#include "../../luxpythonenvgym_b200/csrc/lux_obs_core.cuh"

struct EmuObsArgs {
    const LuxBatch* b; int m; const uint8_t* team_sel; float* obs; int16_t* ent; int32_t* cnt; int e_cap;
    unsigned char* smem; LuxObsLayout L;
};

static void emu_obs_lane(void* p) {
    EmuObsArgs* a = (EmuObsArgs*)p;
    const LuxBatch* b = a->b;
    int m = a->m, ncell = b->size * b->size;
    LuxObsMatch M;
    lux_obs_bind(M, a->smem, a->L, b->size, b->cells + (size_t)m * ncell, b->cities + (size_t)m * b->c_cap);
    lux_copy16(M.hdr, b->hdr + m, LUX_HDR_BYTES);
    lux_sync();
    if (M.hdr->done) { if (lux_lane() == 0) a->cnt[m] = 0; return; }
    lux_copy16(M.units, b->units + (size_t)m * b->u_cap, M.hdr->n_units * 16);
    lux_copy16(M.tiles, b->tiles + (size_t)m * b->t_cap, M.hdr->n_tiles * 2);
    lux_copy16(M.res, b->res + (size_t)m * b->r_cap, M.hdr->n_res * 4);
    lux_sync();
    int n = lux_observe(M, a->team_sel[m], a->obs + (size_t)m * a->e_cap * LUX_OBS_DIM, a->ent + (size_t)m * a->e_cap, a->e_cap);
    if (lux_lane() == 0) a->cnt[m] = n < a->e_cap ? n : a->e_cap;
}

extern "C" int lux_emu_observe(const LuxBatch* b, const uint8_t* team_sel, float* obs, int16_t* ent, int32_t* cnt, int e_cap) {
    if (!b || !team_sel || !obs || !ent || !cnt) return LUX_E_ARG;
    EmuObsArgs a;
    a.b = b; a.team_sel = team_sel; a.obs = obs; a.ent = ent; a.cnt = cnt; a.e_cap = e_cap;
    a.L = lux_obs_layout(b->size, b->u_cap, b->t_cap, b->r_cap);
    a.smem = (unsigned char*)aligned_alloc(16, a.L.total);
    for (int m = 0; m < b->n_matches; m++) {
        memset(a.smem, 0xA5, a.L.total);
        a.m = m;
        warp_emu::run_warp(emu_obs_lane, &a);
    }
    free(a.smem);
    return 0;
}

// ---- AgentPolicy action decoding for one unit (csrc/lux_env_core.cuh) -----------------------
#include "../../luxpythonenvgym_b200/csrc/lux_env_core.cuh"

extern "C" unsigned lux_emu_unit_word(const LuxBatch* b, int m, int slot, int code) {
    const LuxHeader* h = b->hdr + m;
    if (slot < 0 || slot >= h->n_units) return 0;
    return env_unit_word(b->units + (size_t)m * b->u_cap, h->n_units, b->size, slot, code);
}
