// lux_step_core.cuh -- one Lux AI 2021 turn for one match, executed by one lane group (a warp by default).
//
// The match's lists (header, units, cities, tile and resource lists, action words) are staged in the
// group's slice of shared memory, laid out for what this turn can touch (LuxNeed); the cell grid stays in
// global memory and is read in place, a cell at most once per phase; every phase of
// Game.run_turn_with_actions (reference: luxai2021/game/game.py:390-535) runs with lanes striding over
// units / city tiles / resource cells, and the lists are written back in canonical order.
// There is no block-level synchronisation: groups (matches) of a CTA are independent.
//
// This is a parallel re-derivation, not a transcription, of the reference's sequential
// algorithm; oracle/lux_oracle.c keeps the sequential form and the tests compare the two:
//   * spawn-cap validation (game.py:696-718) = per-team ordered prefix via ballots;
//   * collision pruning (game.py:1104-1195)  = static seeds + fixpoint over a per-cell flag plane;
//   * collection (game.py:868-1001)          = per-requested-cell water-filling on a packed
//     histogram of request amounts, then per-worker gather (the iterative equal split is
//     closed-form: request i receives min(amount_i, final level));
//   * order-dependent actions (transfer, build city, pillage; unit.py:162-197) run in unit
//     order on lane 0, everything else is lane-parallel;
//   * dead units / cities are compacted and the canonical orders (team A then B units,
//     cities x city_cells tiles) rebuilt while storing.
#pragma once
#ifndef LUX_EMULATE
__constant__ unsigned lux_phase_sync_mask = 0x54u;       // which of the nine rendezvous points of lux_turn are armed (option "sync_mask";
                                                         // default: before the city-tile turns, the unit pass and the deposit)
#endif
#define LUX_TURN_CTA_SYNC 0x8000u      // internal turn flag: CTA-wide rendezvous between phases (launcher experiment "cta_sync")
#include "../../include/lux_b200.h"
#include "lux_warp.cuh"

// ---- game constants (luxai2021/game/game_constants.json:18-57) -------------------------
#define LUX_DAY_LENGTH 30
#define LUX_CYCLE_LENGTH 40
#define LUX_MAX_DAYS 360
#define LUX_UPKEEP_CITY 23
#define LUX_UPKEEP_WORKER 4
#define LUX_UPKEEP_CART 10
#define LUX_MAX_WOOD 500
#define LUX_CITY_BUILD_COST 100
#define LUX_ADJ_BONUS 5
#define LUX_CAP_WORKER 100
#define LUX_CAP_CART 2000
#define LUX_RESEARCH_COAL 50
#define LUX_RESEARCH_URANIUM 200
#define LUX_CITY_ACTION_COOLDOWN 10
#define LUX_MAX_ROAD_Q 24
#define LUX_ROAD_DEV_Q 3
#define LUX_PILLAGE_Q 2

// per-cell flag plane p1 (movement phase)
#define P1_IN1 0x01u      // >= 1 validated move targets the cell
#define P1_IN2 0x02u      // >= 2
#define P1_STILL 0x04u    // a unit without a validated move stands here
#define P1_BLOCKED 0x08u  // a reverted mover stays here
#define P1_HAS1 0x10u     // >= 1 unit stands here
#define P1_HAS2 0x20u     // >= 2
#define P1_CLAIM0 0x40u   // pre-validation: team 0 / team 1 already holds a validated move into the cell
#define P1_CLAIM1 0x80u

// per-unit scratch byte ust: [2:0] held action kind, [5:3] move direction, flags
#define UST_REVERTED 0x40u
#define UST_FAILED 0x80u  // transfer raised KeyError in the reference: no cooldown, no cart road

// per-unit cache uinfo (built once per turn, after moves and city building): what the unit stands on
#define UI_TILE 0x01000000u   // a city tile; [15:0] = its index in canonical tile order, [23:16] = city slot
#define UI_OWN  0x02000000u   // ... of the unit's own team

struct LuxSmemLayout {
    int p1, occ, mbar, hdr, units, cities, tiles, res, uact, tact, utgt, ust, umine, uinfo, umin, tmask, cprefix, cremap, reqlist, total;
};

LUX_DEV int lux_align16(int x) { return (x + 15) & ~15; }

// What one turn of a match can touch: its live lists plus the most they can grow before the turn ends
// (every city tile may spawn a unit, every unit may found a city with one tile), its resource list and the
// request list of a collection pass.  The shared-memory slice of a match is laid out for exactly this, so a
// typical match (a dozen entities) takes ~3 KB although the batch capacities are 252 / 255 / 384.
struct LuxNeed { int u, c, t, r, q; };

#ifdef LUX_EMULATE
#define LUX_HD static inline
#else
#define LUX_HD __host__ __device__ inline
#endif

LUX_HD LuxNeed lux_need_caps(int u_cap, int c_cap, int t_cap, int r_cap) {
    LuxNeed n; n.u = u_cap; n.c = c_cap; n.t = t_cap; n.r = r_cap; n.q = r_cap;
    return n;
}

// `spawns` = the most units the turn can add: a team spawns only while it has fewer units than city tiles
// (game.py:725-735), and never more than one unit per tile.
LUX_HD LuxNeed lux_need_turn(int n_units, int n_cities, int n_tiles, int n_res, int spawns, int u_cap, int c_cap, int t_cap,
                             int r_cap) {
    LuxNeed n;
    if (spawns > n_tiles) spawns = n_tiles;
    n.u = (n_units + spawns + 3) & ~3;     if (n.u > u_cap) n.u = u_cap;
    n.c = n_cities + n_units;              if (n.c > c_cap) n.c = c_cap;     // new cities: at most one per unit
    if (n.c < 1) n.c = 1;
    n.t = (n_tiles + n_units + 15) & ~15;  if (n.t > t_cap) n.t = t_cap;     // new tiles: at most one per unit
    n.r = (n_res + 7) & ~7;                if (n.r > r_cap) n.r = r_cap;
    n.q = 5 * n.u < n.r ? 5 * n.u : n.r;                                     // requested cells: five per worker at most
    return n;
}
LUX_HD LuxNeed lux_need_none(int u_cap, int c_cap, int t_cap, int r_cap) { return lux_need_turn(0, 0, 0, 0, 0, u_cap, c_cap, t_cap, r_cap); }
// from a header (a header whose per-team counts do not add up gets the safe bound: one spawn per tile)
LUX_HD LuxNeed lux_need_header(const LuxHeader* h, int u_cap, int c_cap, int t_cap, int r_cap) {
    int s0 = (int)h->n_team_tiles[0] - (int)h->n_team_units[0], s1 = (int)h->n_team_tiles[1] - (int)h->n_team_units[1];
    int spawns = (s0 > 0 ? s0 : 0) + (s1 > 0 ? s1 : 0);
    if ((int)h->n_team_tiles[0] + h->n_team_tiles[1] != h->n_tiles || (int)h->n_team_units[0] + h->n_team_units[1] != h->n_units)
        spawns = h->n_tiles;
    return lux_need_turn(h->n_units, h->n_cities, h->n_tiles, h->n_res, spawns, u_cap, c_cap, t_cap, r_cap);
}

// The byte planes, the barrier and the header sit at fixed offsets (they are needed before the match's
// sizes are known); the lists follow in four blocks whose inner offsets are fixed multiples of the (rounded)
// counts -- per-unit arrays, per-tile arrays, resource and request lists, per-city arrays -- so that a warp
// derives the whole layout with a dozen multiply-adds (the first version aligned nineteen fields one by one:
// 9 % of the executed instructions of a turn).  Every bulk-copy destination is 16-byte aligned.
LUX_HD LuxSmemLayout lux_smem_layout_need(int size, LuxNeed n) {
    LuxSmemLayout L;
    const int plane = (size * size + 15) & ~15;
    const int ur = (n.u + 7) & ~7, cr = (n.c + 7) & ~7, tr = (n.t + 15) & ~15, rr = (n.r + 7) & ~7, qr = (n.q + 7) & ~7;
    L.p1 = 0;                           // per-cell flag plane (movement flags, then collection mark / fill level)
    L.occ = plane;                      // per-cell occupant plane (unit index + 1 during collection)
    L.mbar = 2 * plane;
    L.hdr = 2 * plane + 16;
    int o = 2 * plane + 16 + LUX_HDR_BYTES;
    L.units = o; L.uact = o + 16 * ur; L.uinfo = o + 20 * ur; L.utgt = o + 24 * ur; L.umin = o + 26 * ur;
    L.ust = o + 28 * ur; L.umine = o + 29 * ur;
    o += 30 * ur;
    L.tiles = o; L.tact = o + 2 * tr; L.tmask = o + 3 * tr;
    o += 7 * tr;
    L.res = o;
    o += 4 * rr;
    L.reqlist = o;                      // cells requested this collection pass (counter first)
    o += 16 + 2 * qr;
    L.cities = o; L.cprefix = o + 16 * cr; L.cremap = o + 18 * cr + 16;
    o += 19 * cr + 16;
    L.total = (o + 15) & ~15;
    return L;
}

// the fixed head of every slice (same offsets as lux_smem_layout_need gives them)
struct LuxSmemFixed { int p1, occ, mbar, hdr; };
LUX_HD LuxSmemFixed lux_smem_fixed(int size) {
    LuxSmemFixed F;
    const int plane = (size * size + 15) & ~15;
    F.p1 = 0; F.occ = plane; F.mbar = 2 * plane; F.hdr = 2 * plane + 16;
    return F;
}

LUX_HD LuxSmemLayout lux_smem_layout(int size, int u_cap, int c_cap, int t_cap, int r_cap) {
    return lux_smem_layout_need(size, lux_need_caps(u_cap, c_cap, t_cap, r_cap));
}

// The cell grid is NOT staged: `cells` points at the match's grid in global memory and the engine
// touches only the cells its entities stand on, aim at, mine from or own (a few dozen 32-byte
// sectors per turn instead of the whole 8 KB grid).  Everything else lives in shared memory.
struct LuxMatch {
    LuxHeader* hdr;
    LuxCell* cells;      // global memory
    LuxUnit* units;
    LuxCity* cities;
    uint16_t* tiles;
    LuxRes* res;
    LuxRes* g_res;       // the resource list in global memory: the regrowth pass writes the entries it changes straight there
    uint32_t* uact;
    uint8_t* tact;
    uint16_t* utgt;
    uint8_t* ust;
    uint8_t* umine;
    uint8_t* unext;
    uint32_t* uinfo;
    uint16_t* umin;
    uint32_t* tmask;
    uint16_t* cprefix;
    uint8_t* cremap;
    uint16_t* reqlist;
    int* reqcnt;
    uint8_t* p1;
    uint8_t* occ;
    int S, ncell, u_cap, c_cap, t_cap, r_cap;
};

LUX_DEV void lux_match_bind(LuxMatch& M, unsigned char* smem, const LuxSmemLayout& L, LuxCell* g_cells, LuxRes* g_res, int size,
                            const LuxNeed& need) {
    M.hdr = (LuxHeader*)(smem + L.hdr);
    M.cells = g_cells;
    M.g_res = g_res;
    M.units = (LuxUnit*)(smem + L.units);
    M.cities = (LuxCity*)(smem + L.cities);
    M.tiles = (uint16_t*)(smem + L.tiles);
    M.res = (LuxRes*)(smem + L.res);
    M.uact = (uint32_t*)(smem + L.uact);
    M.tact = (uint8_t*)(smem + L.tact);
    M.utgt = (uint16_t*)(smem + L.utgt);
    M.ust = (uint8_t*)(smem + L.ust);
    M.umine = (uint8_t*)(smem + L.umine);
    M.unext = (uint8_t*)(smem + L.utgt);   // next unit (index + 1) on the same non-city cell; shares utgt[], which is done with after the moves
    M.uinfo = (uint32_t*)(smem + L.uinfo);
    M.umin = (uint16_t*)(smem + L.umin);
    M.tmask = (uint32_t*)(smem + L.tmask);
    M.cprefix = (uint16_t*)(smem + L.cprefix);
    M.cremap = (uint8_t*)(smem + L.cremap);
    M.reqcnt = (int*)(smem + L.reqlist);
    M.reqlist = (uint16_t*)(smem + L.reqlist + 16);
    M.p1 = smem + L.p1;
    M.occ = smem + L.occ;
    M.S = size; M.ncell = size * size;
    // the capacities the turn checks overflow against: the batch capacities wherever they bind, otherwise
    // more than the turn can reach
    M.u_cap = need.u; M.c_cap = need.c; M.t_cap = need.t; M.r_cap = need.r;
}

#ifdef LUX_EMULATE
#define LUX_NOINLINE static __attribute__((noinline))
#else
#define LUX_NOINLINE __device__ __noinline__
#endif

// ---- staging with plain 16-byte loads ------------------------------------------------------
// (the CUDA build normally uses cp.async.bulk instead, see lux_step.cu; this form is what the
//  host emulation runs and the device uses when the "tma" option is off)
template <int G = 32>
LUX_DEV void lux_copy16(void* dst, const void* src, int bytes) {   // bytes rounded up to 16
    int lane = lux_lane<G>();
    const uint64_t* s = (const uint64_t*)src;
    uint64_t* d = (uint64_t*)dst;
    int n = (bytes + 15) / 16;
#pragma unroll 1
    for (int i = lane; i < n; i += G) { uint64_t a = s[2 * i], b = s[2 * i + 1]; d[2 * i] = a; d[2 * i + 1] = b; }
}

// ---- cell accessors ----------------------------------------------------------------------
LUX_DEV int cell_is_tile(const LuxCell& c) { return (c.flags & 0x0C) != 0; }
LUX_DEV int cell_tile_team(const LuxCell& c) { return ((c.flags >> 2) & 3) - 1; }
LUX_DEV int cell_res_type(const LuxCell& c) { return c.flags & 3; }
LUX_DEV int cell_adj(const LuxCell& c) { return (c.flags >> 4) & 7; }

// Grid reads go through L1 (the default .ca path): a sector holds four cell records and a unit's neighbourhood
// and the later phases hit the lines the earlier ones fetched; bypassing L1 (ld.global.cg) was measured 18 %
// slower on the bench batch (0.129 against 0.110 ms per step).
#define LUX_LDG64(p) (*(const uint64_t*)(p))
#define LUX_LDG16(p) (*(const uint16_t*)(p))
#define LUX_LDG8(p) (*(const uint8_t*)(p))
// one 8-byte access per cell record
LUX_DEV LuxCell lux_ldcell(const LuxCell* cells, int i) {
    union { uint64_t w; LuxCell c; } v;
    v.w = LUX_LDG64(cells + i);
    return v.c;
}
LUX_DEV void lux_stcell(LuxCell* cells, int i, const LuxCell& c) {
    union { uint64_t w; LuxCell c; } v;
    v.c = c;
    *(uint64_t*)(cells + i) = v.w;
}

LUX_DEV int unit_team(const LuxUnit& u) { return u.team_type & 1; }
LUX_DEV int unit_is_cart(const LuxUnit& u) { return (u.team_type >> 1) & 1; }
LUX_DEV int unit_space(const LuxUnit& u) {   // unit.py:43-52
    return (unit_is_cart(u) ? LUX_CAP_CART : LUX_CAP_WORKER) - ((int)u.wood + (int)u.coal + (int)u.uranium);
}
LUX_DEV int lux_dx(int dir) { return dir == LUX_DIR_EAST ? 1 : dir == LUX_DIR_WEST ? -1 : 0; }
LUX_DEV int lux_dy(int dir) { return dir == LUX_DIR_SOUTH ? 1 : dir == LUX_DIR_NORTH ? -1 : 0; }

// the five cells a worker mines from: own, N, E, S, W (game.py:950-951, game_map.py:484-508)
LUX_DEV int lux_cell5(int S, int x, int y, int k) {
    // branch-free: (dx, dy) + 1 packed two bits per direction, k = 0..4 -> c, n, e, s, w
    int nx = x + (int)((0x065u >> (2 * k)) & 3u) - 1;
    int ny = y + (int)((0x191u >> (2 * k)) & 3u) - 1;
    return ((unsigned)nx < (unsigned)S && (unsigned)ny < (unsigned)S) ? ny * S + nx : -1;
}

template <int G = 32>
LUX_DEV void lux_zero16(void* p, int bytes) {   // bytes multiple of 16, p 16-byte aligned
    int lane = lux_lane<G>();
    uint64_t* q = (uint64_t*)p;
#pragma unroll 1
    for (int i = lane; i < bytes / 16; i += G) { q[2 * i] = 0; q[2 * i + 1] = 0; }
}

// compile-time size: straight 16-byte stores, no loop
template <int G, int BYTES>
LUX_DEV void lux_zero_fixed(void* p) {
    const int lane = lux_lane<G>();
    struct alignas(16) Q { uint64_t a, b; };
    Q* q = (Q*)p;
    Q z; z.a = 0; z.b = 0;
#pragma unroll
    for (int i = 0; i < (BYTES / 16 + G - 1) / G; i++)
        if (i * G + lane < BYTES / 16) q[i * G + lane] = z;
}

// umin[] holds 2 bits of resource type for each of a worker's five cells; bit 2j of the result is set where
// field j equals `rtype` (1..3)
LUX_DEV unsigned lux_fields_equal(unsigned mm, int rtype) {
    unsigned x = mm ^ ((unsigned)rtype * 0x155u);
    return ~(x | (x >> 1)) & 0x155u;
}

// The two per-cell byte planes must be all-zero when a turn starts; the engine clears every byte it
// sets before the turn ends, so a warp zeroes them once (here) and then plays any number of matches.
template <int G = 32>
LUX_DEV void lux_planes_init(LuxMatch& M) {
    lux_zero16<G>(M.p1, lux_align16(M.ncell));
    lux_zero16<G>(M.occ, lux_align16(M.ncell));
}

// A group without a match to play (past the end of the batch, finished, does not fit this launch's slice)
// walks through the turn in lockstep with the other groups of its warp on an empty match: all counts zero.
template <int G = 32>
LUX_DEV void lux_idle_header(LuxMatch& M) { lux_zero16<G>(M.hdr, LUX_HDR_BYTES); }

// ---- staging with plain loads: header first, then the lists at the sizes the header names --------------
// Pointers into one match's sections in global memory.
struct LuxGlobalMatch {
    LuxHeader* hdr; LuxCell* cells; LuxUnit* units; LuxCity* cities; uint16_t* tiles; LuxRes* res;
    const uint32_t* uact; const uint8_t* tact;
};
enum { LUX_M_DONE = 0, LUX_M_SPILL = 1, LUX_M_BAD = 2 };     // why a group is not playing: nothing to play / the turn does not fit the slice / unplayable header

// Returns the group's `active` flag (0: nothing to play -- invalid, finished, or the turn's footprint does not
// fit `slice` bytes, in which case *result says which) with M bound and, when active, staged.  An inactive
// group is bound to an empty match.  Warp-uniform control flow: every lane of the warp must call this.
template <int G>
LUX_DEV int lux_stage_plain(LuxMatch& M, unsigned char* smem, const LuxGlobalMatch& g, int valid, int size, int u_cap, int c_cap,
                            int t_cap, int r_cap, int slice, int* result) {
    const LuxSmemFixed L0 = lux_smem_fixed(size);
    LuxHeader* sh = (LuxHeader*)(smem + L0.hdr);
    int active = valid;
    *result = LUX_M_DONE;
    if (active) lux_copy16<G>(sh, g.hdr, LUX_HDR_BYTES);
    lux_sync<G>();
    if (active && sh->done) active = 0;
    // live counts beyond the capacities, or per-team unit counts that do not add up: not playable (the caller flags it)
    if (active && (sh->n_units > u_cap || sh->n_cities > c_cap || sh->n_tiles > t_cap || sh->n_res > r_cap ||
                   (int)sh->n_team_units[0] + (int)sh->n_team_units[1] != (int)sh->n_units)) {
        active = 0; *result = LUX_M_BAD;
    }
    LuxNeed need = lux_need_none(u_cap, c_cap, t_cap, r_cap);
    if (active) need = lux_need_header(sh, u_cap, c_cap, t_cap, r_cap);
    LuxSmemLayout L = lux_smem_layout_need(size, need);
    if (active && L.total > slice) {
        active = 0; *result = LUX_M_SPILL;
        need = lux_need_none(u_cap, c_cap, t_cap, r_cap);
        L = lux_smem_layout_need(size, need);
    }
    lux_match_bind(M, smem, L, g.cells, g.res, size, need);
    lux_sync<G>();                            // every lane has read the header before an idle group clears it
    if (active) {
        lux_copy16<G>(M.units, g.units, M.hdr->n_units * 16);
        lux_copy16<G>(M.cities, g.cities, M.hdr->n_cities * 16);
        lux_copy16<G>(M.tiles, g.tiles, M.hdr->n_tiles * 2);
        lux_copy16<G>(M.res, g.res, M.hdr->n_res * 4);
        lux_copy16<G>(M.uact, g.uact, M.hdr->n_units * 4);
        lux_copy16<G>(M.tact, g.tact, M.hdr->n_tiles);
    } else {
        lux_idle_header<G>(M);
    }
    lux_sync<G>();
    return active;
}

// ---- collection of one resource type (game.py:882-1001) -----------------------------------
// One inlined copy inside a rolled loop over the resource types, inner loops rolled: the kernel is
// instruction-fetch sensitive (flat, once-through code far larger than the instruction caches).
// Works on shared memory only, except for the amount of each requested cell (one global load and one
// store per cell): which neighbours a worker can mine was cached in umin[] by the unit pass.
template <int kS, int G>
LUX_DEV void lux_collect_type(LuxMatch& M, int rtype, int n_units, int researched0, int researched1, int stacked) {
    const int lane = lux_lane<G>();
    const int S = kS > 0 ? kS : M.S;
    LuxHeader* h = M.hdr;
    const int rate = rtype == LUX_RES_WOOD ? 20 : rtype == LUX_RES_COAL ? 5 : 2;
    const int fuel_rate = rtype == LUX_RES_WOOD ? 1 : rtype == LUX_RES_COAL ? 10 : 40;

    if (lane == 0) *M.reqcnt = 0;
    lux_sync<G>();

    // W1: every eligible worker computes its request amount (create_resource_requests :933-967)
    // and registers the cells it mines from in the request list
#pragma unroll 1
    for (int u = lane; u < n_units; u += G) {
        const LuxUnit& U = M.units[u];
        int mine = 0xFF;
        unsigned z = lux_fields_equal(M.umin[u], rtype);     // the cells of this type it can mine (none for carts)
        if (z && (unit_team(U) ? researched1 : researched0)) {
            const int k = lux_popc(z);
#pragma unroll 1
            while (z) {
                int j = (lux_ffs(z) - 1) >> 1;
                z &= z - 1;
                int c = lux_cell5(S, U.x, U.y, j);
                unsigned old = lux_atomic_or_byte(M.p1, c, 0x80u);
                if (!(old & 0x80u)) M.reqlist[lux_atomic_add(M.reqcnt, 1)] = (uint16_t)c;
            }
            {
                int space = unit_space(U);
                // ceil(space / k), k = 1..5, space <= 2000: multiply by ceil(2^16 / k) instead of dividing
                unsigned inv = k == 1 ? 65536u : k == 2 ? 32768u : k == 3 ? 21846u : k == 4 ? 16384u : 13108u;
                mine = (int)(((unsigned)(space + k - 1) * inv) >> 16);
                if (mine > rate) mine = rate;
                unsigned info = M.uinfo[u];
                if (info & UI_TILE) {
                    // requests from one city tile are de-duplicated by amount (:920-931, :964-966); the
                    // occupant plane names any one requester standing on the tile
                    lux_atomic_or(&M.tmask[info & 0xFFFFu], 1u << mine);
                    M.occ[U.y * S + U.x] = (uint8_t)(u + 1);
                }
            }
        }
        M.umine[u] = (uint8_t)mine;
    }
    lux_sync<G>();

    // C: every requested cell resolves its requests (resolve_resource_requests :969-1001)
    const int n_req = *M.reqcnt;
#pragma unroll 1
    for (int q = lane; q < n_req; q += G) {
        int c = M.reqlist[q];
        int left = LUX_LDG16(&M.cells[c].amount);       // global; consumed after the gather below
        int level = 0;
        int x = c % S, y = c / S;
        // histogram of request amounts 0..20, 3 bits each (at most one request per amount and source cell)
        uint64_t hist = 0;
        unsigned present = 0;                         // amounts that occur
        unsigned tile_dirs = 0;                       // neighbours that are city tiles holding requests
        int n = 0, sum = 0;
        // the occupants of the five source cells, one byte each (most are empty)
        uint64_t occ5 = 0;
        int crowded = 0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            int p = lux_cell5(S, x, y, j);
            if (p >= 0) occ5 |= (uint64_t)M.occ[p] << (8 * j);
        }
#pragma unroll 1
        while (occ5) {
            int j = ((occ5 & 0xFFFFFFFFull) ? lux_ffs((unsigned)occ5) - 1 : 32 + lux_ffs((unsigned)(occ5 >> 32)) - 1) >> 3;
            int o = (int)((occ5 >> (8 * j)) & 0xFFu);
            occ5 &= ~((uint64_t)0xFF << (8 * j));
            unsigned info = M.uinfo[o - 1];
            if (info & UI_TILE) {
                unsigned mask = M.tmask[info & 0xFFFFu];
                if (mask) tile_dirs |= 1u << j;
                while (mask) {
                    int a = lux_ffs(mask) - 1;
                    mask &= mask - 1;
                    hist += (uint64_t)1 << (3 * a);
                    present |= 1u << a;
                    n++; sum += a;
                }
            } else {
                int a = M.umine[o - 1];
                if (stacked) {
                    // further units stacked on the same non-city cell (left behind by a fallen city) follow in
                    // a chain; the histogram counts up to seven equal requests
                    if (a != 0xFF && ((hist >> (3 * a)) & 7) == 7) crowded = 1;
                    occ5 |= (uint64_t)M.unext[o - 1] << (8 * j);
                }
                if (a != 0xFF) { hist += (uint64_t)1 << (3 * a); present |= 1u << a; n++; sum += a; }
            }
        }
        if (n > 0 && sum > 0) {
#pragma unroll 1
            while (present && n > 0 && left > 0) {        // ascending over the amounts that occur
                int a = lux_ffs(present) - 1;
                present &= present - 1;
                int cnt = (int)((hist >> (3 * a)) & 7);
                int minrem = a - level;
                int qq = left / n;
                if (qq >= minrem) {
                    left -= minrem * n;
                    level = a;
                    if (left < n) left = 0;
                    n -= cnt;
                } else {
                    level += qq;
                    left = 0;
                }
            }
            M.cells[c].amount = (uint16_t)left;
            // city-tile requests are credited here (game.py:988-990)
            if (level > 0) {
#pragma unroll 1
                while (tile_dirs) {
                    int j = lux_ffs(tile_dirs) - 1;
                    tile_dirs &= tile_dirs - 1;
                    int p = lux_cell5(S, x, y, j);
                    unsigned info = M.uinfo[M.occ[p] - 1];
                    unsigned mask = M.tmask[info & 0xFFFFu];
                    int got = 0;
                    while (mask) {
                        int a = lux_ffs(mask) - 1;
                        mask &= mask - 1;
                        got += a < level ? a : level;
                    }
                    if (got) {
                        LuxCity& city = M.cities[(info >> 16) & 0xFFu];
                        lux_atomic_add(&city.fuel, got * fuel_rate);
                        lux_atomic_add(&h->collected[city.team][rtype - 1], got);
                    }
                }
            }
        }
        M.p1[c] = (uint8_t)(0x80u | level);           // fill level; the request mark stays for the regrowth pass
        if (crowded) h->status |= LUX_ST_BAD_STATE;
    }
    lux_sync<G>();

    // W2: workers off city tiles take min(request, level) from each cell, capped by cargo space
    int got0 = 0, got1 = 0;
#pragma unroll 1
    for (int u = lane; u < n_units; u += G) {
        int mine = M.umine[u];
        if (mine == 0xFF) continue;
        unsigned info = M.uinfo[u];
        if (info & UI_TILE) {
            M.tmask[info & 0xFFFFu] = 0;                   // leave the mask plane clean for the next type
            continue;
        }
        LuxUnit& U = M.units[u];
        unsigned z = lux_fields_equal(M.umin[u], rtype);
        int gain = 0;
#pragma unroll 1
        while (z) {
            int j = (lux_ffs(z) - 1) >> 1;
            z &= z - 1;
            int lv = M.p1[lux_cell5(S, U.x, U.y, j)] & 0x7F;
            gain += mine < lv ? mine : lv;
        }
        int space = unit_space(U);
        if (gain > space) gain = space;
        if (rtype == LUX_RES_WOOD) U.wood = (uint16_t)(U.wood + gain);
        else if (rtype == LUX_RES_COAL) U.coal = (uint16_t)(U.coal + gain);
        else U.uranium = (uint16_t)(U.uranium + gain);
        if (unit_team(U)) got1 += gain; else got0 += gain;
    }
    {   // one reduction: a team gains at most 5 cells x 20 x 252 workers < 2^16 per pass
        unsigned packed = (unsigned)lux_reduce_add<G>(got0 | (got1 << 16));
        got0 = (int)(packed & 0xFFFFu); got1 = (int)(packed >> 16);
    }
    if (lane == 0) {
        h->collected[0][rtype - 1] += got0;
        h->collected[1][rtype - 1] += got1;
    }
    // the levels of this pass stay in p1 until the regrowth pass clears them: the next type's cells are
    // disjoint from these (a cell holds one resource type)
    lux_sync<G>();
}

// exclusive prefix of city tile counts -> M.cprefix[0..n_cities]
template <int G = 32>
LUX_DEV void lux_city_prefix(LuxMatch& M, int n_cities) {
    const int lane = lux_lane<G>();
    const int nc_max = lux_wmax<G>(n_cities);
    int run = 0;
#pragma unroll 1
    for (int base = 0; base < nc_max; base += G) {
        int s = base + lane;
        int v = s < n_cities ? M.cities[s].ntiles : 0;
        int tot;
        int ex = lux_exscan_add<G>(v, &tot);
        if (s < n_cities) M.cprefix[s] = (uint16_t)(run + ex);
        run += tot;
    }
    if (lane == 0) M.cprefix[n_cities] = (uint16_t)run;
    lux_sync<G>();
}

// utgt[u]: [9:0] target cell of the unit's validated move (own cell otherwise), bit 14 = the target is a city tile
#define UTGT_TILE 0x4000u
#define UTGT_CELL 0x03FFu

// Units that precede unit `u` in unit order and share its cell (a stack; movers count where they arrive) have had
// their turn on the open cell before `u` founds a city there: a cart among them has added its road step, which the
// unit pass -- it finds a city tile -- will not do.  Replayed in unit order (pillages of the cell included: the first
// pillager recorded the level the turn opened with); returns the cell's road level.  Lane 0 only.
template <int kS>
LUX_DEV int lux_road_before_city(LuxMatch& M, int u, int idx, int road_now) {
    LuxHeader* h = M.hdr;
    const int S = kS > 0 ? kS : M.S;
    int opened = -1, carts = 0;
#pragma unroll 1
    for (int v = 0; v < u; v++) {
        const unsigned sv = M.ust[v];
        const LuxUnit& V = M.units[v];
        const int at = (sv & UST_REVERTED) ? V.y * S + V.x : (int)(M.utgt[v] & UTGT_CELL);
        if (at != idx) continue;
        if (unit_is_cart(V)) carts++;
        else if ((sv & 7) == LUX_UA_PILLAGE && opened < 0) opened = M.umine[v];
    }
    if (!carts) return road_now;
    int road = opened >= 0 ? opened : road_now;
#pragma unroll 1
    for (int v = 0; v < u; v++) {
        const unsigned sv = M.ust[v];
        const LuxUnit& V = M.units[v];
        const int at = (sv & UST_REVERTED) ? V.y * S + V.x : (int)(M.utgt[v] & UTGT_CELL);
        if (at != idx) continue;
        if (unit_is_cart(V)) {
            if ((sv & UST_FAILED) || road >= LUX_MAX_ROAD_Q) continue;
            road = road + LUX_ROAD_DEV_Q > LUX_MAX_ROAD_Q ? LUX_MAX_ROAD_Q : road + LUX_ROAD_DEV_Q;
            h->roads_built_q[unit_team(V)] += LUX_ROAD_DEV_Q;
        } else if ((sv & 7) == LUX_UA_PILLAGE) {
            road = road > LUX_PILLAGE_Q ? road - LUX_PILLAGE_Q : 0;
        }
    }
    return road;
}

// ---- build a city tile under unit `u` (spawn_city_tile game.py:798-854); lane 0 only ---------
template <int kS, int G>
LUX_DEV void lux_build_city_lane0(LuxMatch& M, int u, int& n_cities, int& n_tiles, int prevalidate) {
    LuxHeader* h = M.hdr;
    LuxUnit& U = M.units[u];
    const int S = kS > 0 ? kS : M.S;
    const int team = unit_team(U);
    const int idx = U.y * S + U.x;
    // the cell and its four neighbours: five independent global loads
    int pos[5];
    LuxCell nb[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        pos[j] = lux_cell5(S, U.x, U.y, j);
        if (pos[j] >= 0) nb[j] = lux_ldcell(M.cells, pos[j]);
    }
    LuxCell C = nb[0];
    int same[4], nsame = 0, found[4], nfound = 0;
#pragma unroll
    for (int j = 1; j < 5; j++) {
        if (pos[j] < 0) continue;
        const LuxCell& P = nb[j];
        if (cell_is_tile(P) && cell_tile_team(P) == team) {
            same[nsame++] = j;
            int seen = 0;
            for (int f = 0; f < nfound; f++) seen |= (found[f] == P.city_slot);
            if (!seen) found[nfound++] = P.city_slot;
        }
    }
    // A second unit of a stack founding a city on the cell an earlier one has just turned into a tile: the reference
    // builds a second CityTile object over the first and leaves the cell in two cities' lists (spawn_city_tile does not
    // look at the cell itself, game.py:798-854) -- not a state the packed form can hold.  Flagged, nothing built.
    if (cell_is_tile(C)) { h->status |= LUX_ST_BAD_STATE; return; }
    if (n_tiles >= M.t_cap) { h->status |= LUX_ST_TILE_OVERFLOW; return; }
    if (nsame == 0 && n_cities >= M.c_cap) { h->status |= LUX_ST_CITY_OVERFLOW; return; }
    // the cell turns into a city tile NOW, in this unit's turn: what earlier units of a stack did on the open cell first.
    // The movement flags of the cell are still up (second occupant / a validated move aiming at it); with
    // pre-validation they are only set when some unit of the warp moves, so that mode always looks.
    if (u > 0 && (prevalidate || (M.p1[idx] & (P1_HAS2 | P1_IN1)))) C.road_q = (uint8_t)lux_road_before_city<kS>(M, u, idx, C.road_q);
    if (nsame == 0) {
        LuxCity& city = M.cities[n_cities];
        city.id = ++h->city_id_count;
        city.fuel = 0; city.ntiles = 1; city.adjsum = 0; city.team = (uint8_t)team;
        city.pad[0] = city.pad[1] = city.pad[2] = 0;
        C.flags = (uint8_t)((C.flags & 0x03) | ((team + 1) << 2));
        C.tile_cd = 0; C.city_slot = (uint8_t)n_cities; C.key = 0;
        lux_stcell(M.cells, idx, C);
        M.tiles[n_tiles++] = (uint16_t)idx;
        n_cities++;
        return;
    }
    const int cid = found[0];
    LuxCity& city = M.cities[cid];
    C.flags = (uint8_t)((C.flags & 0x03) | ((team + 1) << 2) | (nsame << 4));
    C.tile_cd = 0; C.city_slot = (uint8_t)cid; C.key = city.ntiles;
    lux_stcell(M.cells, idx, C);
    for (int k = 0; k < nsame; k++) M.cells[pos[same[k]]].flags = (uint8_t)(nb[same[k]].flags + 0x10);
    city.ntiles = (uint16_t)(city.ntiles + 1);
    city.adjsum = (uint16_t)(city.adjsum + 2 * nsame);
    M.tiles[n_tiles++] = (uint16_t)idx;
    for (int f = 1; f < nfound; f++) {      // merge the other adjacent cities (:844-852)
        LuxCity& old = M.cities[found[f]];
        int base = city.ntiles;
#pragma unroll 1
        for (int t = 0; t < n_tiles; t++) {
            int tc = M.tiles[t];
            LuxCell T = lux_ldcell(M.cells, tc);
            if (cell_is_tile(T) && T.city_slot == found[f]) {
                T.city_slot = (uint8_t)cid;
                T.key = (uint16_t)(T.key + base);
                lux_stcell(M.cells, tc, T);
            }
        }
        city.ntiles = (uint16_t)(city.ntiles + old.ntiles);
        city.adjsum = (uint16_t)(city.adjsum + old.adjsum);
        city.fuel += old.fuel;
        old.ntiles = 0; old.adjsum = 0; old.fuel = 0;
    }
}

// Several units share a non-city cell (game.py:1173: a city fell under them and they survived the night): the OR-ed
// occupant bytes of those cells are meaningless; the occupants are rebuilt as chains occ[cell] -> unit -> unext -> ...
// in unit order, and the road work of each such cell is replayed in unit order.  Lane 0 only; rare.
template <int kS>
LUX_DEV void lux_stack_paths(LuxMatch& M, int n_units) {
    LuxHeader* h = M.hdr;
    const int S = kS > 0 ? kS : M.S;
#pragma unroll 1
    for (int u = 0; u < n_units; u++) {
        M.unext[u] = 0;
        if (!(M.uinfo[u] & UI_TILE)) M.occ[M.units[u].y * S + M.units[u].x] = 0;
    }
#pragma unroll 1
    for (int u = n_units - 1; u >= 0; u--) {
        if (M.uinfo[u] & UI_TILE) continue;
        const int cell = M.units[u].y * S + M.units[u].x;
        M.unext[u] = M.occ[cell];
        M.occ[cell] = (uint8_t)(u + 1);
    }
    // The road of a stack's cell is worked on IN TURN (Worker.turn / Cart.turn, unit.py:188-192, :257-269):
    // unit after unit, a pillaging worker takes its step off, a cart adds its own while the road is below
    // the maximum.  The engine has applied the pillages first (ordered section) and then let every cart's
    // lane add a step to the level it loaded; for a cell that holds two or more such units the sequence
    // is replayed here in unit order from the level the first of them found, and the statistics
    // corrected by the difference.
#pragma unroll 1
    for (int u = 0; u < n_units; u++) {
        if ((M.uinfo[u] & UI_TILE) || !M.unext[u]) continue;
        const int cell = M.units[u].y * S + M.units[u].x;
        if (M.occ[cell] != u + 1) continue;                      // not the head of its chain
        int nact = 0, r0 = -1, c0 = 255, par0 = 0, par1 = 0;
#pragma unroll 1
        for (int v = u + 1; v; v = M.unext[v - 1]) {
            const LuxUnit& V = M.units[v - 1];
            if (unit_is_cart(V)) {
                const int seen = M.umine[v - 1] & 0x7F, failed = M.umine[v - 1] & 0x80;
                nact++;
                if (seen < c0) c0 = seen;
                if (!failed && seen < LUX_MAX_ROAD_Q) { if (unit_team(V)) par1 += LUX_ROAD_DEV_Q; else par0 += LUX_ROAD_DEV_Q; }
            } else if ((M.ust[v - 1] & 7) == LUX_UA_PILLAGE) {
                nact++;
                if (r0 < 0) r0 = M.umine[v - 1];                 // the first pillager saw the turn's opening level
            }
        }
        if (nact < 2) continue;                                  // a single step was applied as it should be
        int road = r0 >= 0 ? r0 : c0, seq0 = 0, seq1 = 0;
#pragma unroll 1
        for (int v = u + 1; v; v = M.unext[v - 1]) {
            const LuxUnit& V = M.units[v - 1];
            if (unit_is_cart(V)) {
                if ((M.umine[v - 1] & 0x80) || road >= LUX_MAX_ROAD_Q) continue;
                road = road + LUX_ROAD_DEV_Q > LUX_MAX_ROAD_Q ? LUX_MAX_ROAD_Q : road + LUX_ROAD_DEV_Q;
                if (unit_team(V)) seq1 += LUX_ROAD_DEV_Q; else seq0 += LUX_ROAD_DEV_Q;
            } else if ((M.ust[v - 1] & 7) == LUX_UA_PILLAGE) {
                road = road > LUX_PILLAGE_Q ? road - LUX_PILLAGE_Q : 0;
            }
        }
        M.cells[cell].road_q = (uint8_t)road;
        h->roads_built_q[0] += seq0 - par0;
        h->roads_built_q[1] += seq1 - par1;
    }
}

// ---- the turn ------------------------------------------------------------------------------
// On entry the match's lists are staged in shared memory (M.hdr, units[0..n_units), cities, tiles,
// res, uact, tact), the two byte planes are all-zero, M.cells points at the grid in global memory
// and hdr.done == 0.  On exit the grid holds the post-turn cells, shared memory the post-turn
// header / cities (uncompacted) / units (uncompacted, ust says who died); lux_finish_and_store
// compacts while writing.
struct LuxTurnOut { int n_units; int n_cities; int n_tiles; int tiles_dirty; int alive0; int alive1; int tt0; int tt1; };


template <int kS, int G>
LUX_DEV LuxTurnOut lux_turn(LuxMatch& M, unsigned flags) {
    // Phase rendezvous of the CTA's warps (LUX_TURN_CTA_SYNC, set by the launcher for dense batches): warps that enter
    // a phase together fetch its code together (the turn is a once-through program larger than the SM's 32 KB
    // instruction cache).  Every warp of the CTA walks through lux_turn (idle ones on an empty match), so every warp
    // executes the same barriers.
#ifdef LUX_EMULATE
#define LUX_PHASE_SYNC(i) do { } while (0)
#else
#define LUX_PHASE_SYNC(i) do { if ((flags & LUX_TURN_CTA_SYNC) && ((lux_phase_sync_mask >> (i)) & 1u)) asm volatile("bar.sync 1, %0;" :: "r"(blockDim.x) : "memory"); } while (0)
#endif
    const int lane = lux_lane<G>();
    const unsigned lt = lux_lanemask_lt<G>();
    LuxHeader* h = M.hdr;
    const int S = kS > 0 ? kS : M.S;
    const int n_units0 = h->n_units;
    const int n_tiles0 = h->n_tiles;
    int n_cities = h->n_cities;
    int n_tiles = n_tiles0;
    const int turn = h->turn;
    const int night = (turn % LUX_CYCLE_LENGTH) >= LUX_DAY_LENGTH;   // game.py:1197-1204
    const int mult = night ? 2 : 1;
    int tiles_dirty = 0;

    {
        int nt = n_tiles0 + n_units0;                 // tile indices that can exist during this turn
        if (nt > M.t_cap) nt = M.t_cap;
        lux_zero16<G>(M.tmask, lux_align16(nt * 4));
    }

    // ---- team city-tile counts (worker_unit_cap_reached game.py:725-735)
    // the header carries them (lux_finish_and_store, the host map generator); a header whose per-team counts
    // do not add up is recounted from the cities
    int tt0 = h->n_team_tiles[0], tt1 = h->n_team_tiles[1];
    if (lux_wany<G>(tt0 + tt1 != n_tiles0)) {
        int c0 = 0, c1 = 0;
#pragma unroll 1
        for (int s = lane; s < n_cities; s += G) {
            if (M.cities[s].team) c1 += M.cities[s].ntiles; else c0 += M.cities[s].ntiles;
        }
        c0 = lux_reduce_add<G>(c0);
        c1 = lux_reduce_add<G>(c1);
        if (tt0 + tt1 != n_tiles0) { tt0 = c0; tt1 = c1; }
    }
    const int tu0 = h->n_team_units[0], tu1 = h->n_team_units[1];

    int any_move = 0;
    LUX_PHASE_SYNC(0);
    // ---- unit action validation (validate_command game.py:646-695), lanes over units; at most one
    // global cell load per unit (its own cell for a build, the target cell for a move)
#pragma unroll 1
    for (int u = lane; u < n_units0; u += G) {
        const LuxUnit& U = M.units[u];
        unsigned w = M.uact[u];
        int kind = w & 7;
        int ucell = U.y * S + U.x;
        unsigned st = 0;
        unsigned tgt = ucell;
        if (kind == LUX_UA_BCITY) {
            if (U.cd_q < 4 && (int)U.wood + (int)U.coal + (int)U.uranium >= LUX_CITY_BUILD_COST) {
                LuxCell C = lux_ldcell(M.cells, ucell);
                if (!cell_is_tile(C) && C.amount == 0) st = kind;
            }
        } else if (kind == LUX_UA_MOVE) {
            int dir = (w >> 3) & 7;
            if (U.cd_q < 4 && dir <= 4) {
                int ok = 1;
                if (dir != LUX_DIR_CENTER) {
                    int nx = U.x + lux_dx(dir), ny = U.y + lux_dy(dir);
                    if (nx < 0 || ny < 0 || nx >= S || ny >= S) ok = 0;
                    else {
                        tgt = ny * S + nx;
                        LuxCell T = lux_ldcell(M.cells, tgt);
                        if (cell_is_tile(T)) {
                            if (cell_tile_team(T) != unit_team(U)) ok = 0;
                            tgt |= UTGT_TILE;
                        }
                    }
                } else {
                    // a CENTER move targets the unit's own cell: its tile-ness decides whether it collides
                    LuxCell C = lux_ldcell(M.cells, ucell);
                    if (cell_is_tile(C)) tgt |= UTGT_TILE;
                }
                if (ok) st = kind | (dir << 3);
                else tgt = ucell;
            }
        } else if (kind == LUX_UA_PILLAGE || kind == LUX_UA_TRANSFER) {
            st = kind;                                             // game.py:721-723: no check
            if (flags & LUX_STEP_PREVALIDATE) {
                // Action.is_valid (actions.py:322-345, :368-385): the acting unit can act; a transfer
                // names another unit of the team standing at Manhattan distance <= 1
                if (!(U.cd_q < 4)) st = 0;
                if (st && kind == LUX_UA_TRANSFER) {
                    int dest_id = (int)(w >> 17), found = 0;
                    for (int v = 0; v < n_units0; v++) {
                        const LuxUnit& V = M.units[v];
                        if (V.id == dest_id && unit_team(V) == unit_team(U) && v != u) {
                            int d = (V.x > U.x ? V.x - U.x : U.x - V.x) + (V.y > U.y ? V.y - U.y : U.y - V.y);
                            found = d <= 1;
                            break;
                        }
                    }
                    if (!found) st = 0;
                }
            }
        }
        M.ust[u] = (uint8_t)st;
        M.utgt[u] = (uint16_t)tgt;
        any_move |= (st & 7) == LUX_UA_MOVE;
        if (!(flags & LUX_STEP_PREVALIDATE)) {
            // the movement flags of this unit (see the collision stage below) go out in the same pass; with
            // pre-validation they depend on which moves survive it and are set after that stage
            const int moving = (st & 7) == LUX_UA_MOVE;
            unsigned old = lux_atomic_or_byte(M.p1, ucell, P1_HAS1 | (moving ? 0u : P1_STILL));
            if (old & P1_HAS1) lux_atomic_or_byte(M.p1, ucell, P1_HAS2);
            if (moving && !(tgt & UTGT_TILE)) {
                unsigned o2 = lux_atomic_or_byte(M.p1, tgt & UTGT_CELL, P1_IN1);
                if (o2 & P1_IN1) lux_atomic_or_byte(M.p1, tgt & UTGT_CELL, P1_IN2);
            }
        }
    }
    lux_sync<G>();
    // no validated move anywhere in the warp: the whole collision stage is a no-op (a group without moves
    // that walks through it with its neighbours sets and clears a few flags, nothing else)
    const int any_move_w = lux_any_warp(any_move);
    const int nu0_max = lux_wmax<G>(n_units0);

    // ---- MatchController pre-validation of moves (MoveAction.is_valid actions.py:58-116): of the
    // same-team moves into one non-city cell only the first in unit order is buffered; a CENTER
    // move is the do-nothing action and never changes the outcome, so it is dropped here
    if ((flags & LUX_STEP_PREVALIDATE) && any_move_w) {
#pragma unroll 1
        for (int base = 0; base < nu0_max; base += G) {
            int u = base + lane;
            unsigned st = u < n_units0 ? M.ust[u] : 0;
            int is_move = (st & 7) == LUX_UA_MOVE;
            int center = is_move && ((st >> 3) & 7) == LUX_DIR_CENTER;
            unsigned tg = is_move ? M.utgt[u] : 0;
            int t = tg & UTGT_CELL;
            int team = is_move ? unit_team(M.units[u]) : 0;
            int cand = is_move && !center && !(tg & UTGT_TILE);
            unsigned peers = lux_match_any<G>(cand ? ((t << 1) | team) : (int)(0x1000 | lane));
            if (center) M.ust[u] = 0;
            if (cand) {
                unsigned bit = team ? P1_CLAIM1 : P1_CLAIM0;
                int first = lux_ffs(peers) - 1 == lane;
                if (!first || (M.p1[t] & bit)) M.ust[u] = 0;
            }
            lux_sync<G>();
            if (cand && M.ust[u]) lux_atomic_or_byte(M.p1, t, team ? P1_CLAIM1 : P1_CLAIM0);
            lux_sync<G>();
        }
    }

    LUX_PHASE_SYNC(1);
    // ---- collision pruning (handle_movement_actions game.py:1104-1195); shared memory only.
    // Flags are set on every cell (city tiles too) but only ever read on non-city target cells.
    if (any_move_w) {
    if (flags & LUX_STEP_PREVALIDATE) {
#pragma unroll 1
    for (int u = lane; u < n_units0; u += G) {
        const LuxUnit& U = M.units[u];
        int ucell = U.y * S + U.x;
        unsigned st = M.ust[u];
        int moving = (st & 7) == LUX_UA_MOVE;
        unsigned old = lux_atomic_or_byte(M.p1, ucell, P1_HAS1 | (moving ? 0u : P1_STILL));
        if (old & P1_HAS1) lux_atomic_or_byte(M.p1, ucell, P1_HAS2);
        if (moving) {
            unsigned tg = M.utgt[u];
            if (!(tg & UTGT_TILE)) {
                int t = tg & UTGT_CELL;
                unsigned o2 = lux_atomic_or_byte(M.p1, t, P1_IN1);
                if (o2 & P1_IN1) lux_atomic_or_byte(M.p1, t, P1_IN2);
            }
        }
    }
    lux_sync<G>();
    }
    // One monotone closure, iterated to its fixpoint: a move is reverted when >= 2 moves aim at its (non-city)
    // target cell, or the cell's single occupant is not moving (the seeds, :1164-1179), or a reverted mover stays
    // there (the cascade, revert_action :1138-1156: a reverted mover blocks its own non-city cell).  Whatever order
    // the lanes see each other's BLOCKED flags in, the fixpoint is the set the reference's recursion computes; a
    // turn without collisions (the common one) costs a single pass.
#pragma unroll 1
    for (;;) {
        int changed = 0;
#pragma unroll 1
        for (int u = lane; u < n_units0; u += G) {
            unsigned st = M.ust[u];
            if ((st & 7) != LUX_UA_MOVE || (st & UST_REVERTED)) continue;
            unsigned tg = M.utgt[u];
            if (tg & UTGT_TILE) continue;
            const unsigned f = M.p1[tg & UTGT_CELL];
            if ((f & (P1_BLOCKED | P1_IN2)) || ((f & P1_HAS1) && !(f & P1_HAS2) && (f & P1_STILL))) {
                M.ust[u] = (uint8_t)(st | UST_REVERTED);
                const LuxUnit& U = M.units[u];
                lux_atomic_or_byte(M.p1, U.y * S + U.x, P1_BLOCKED);
                changed = 1;
            }
        }
        lux_sync<G>();
        if (!lux_any_warp(changed)) break;
    }
    }   // any_move_w
    // (the movement flags are cleared at the head of the unit pass: the byte of every cell a unit stands on or aimed at)

    LUX_PHASE_SYNC(2);
    // ---- city tiles: spawn validation (game.py:696-718) + CityTile.turn (city.py:96-120)
    int n_units = n_units0;
    int born0 = 0, born1 = 0, died0 = 0, died1 = 0;      // per-team population changes of this turn
    {
        int acc0 = 0, acc1 = 0, spawned = 0, rp0 = 0, rp1 = 0;
        int wb0 = 0, wb1 = 0, cb0 = 0, cb1 = 0;
        const int uid0 = h->unit_id_count;
        const int room = M.u_cap - n_units0;
        const int nt0_max = lux_wmax<G>(n_tiles0);
#pragma unroll 1
        for (int base = 0; base < nt0_max; base += G) {
            int p = base + lane;
            int valid = p < n_tiles0;
            int code = valid ? M.tact[p] : 0;
            int cell = valid ? M.tiles[p] : 0;
            LuxCell C;
            C.flags = 0; C.tile_cd = 0;
            if (valid) C = lux_ldcell(M.cells, cell);
            int team = valid ? cell_tile_team(C) : 0;
            int cd = C.tile_cd;
            if (!lux_wany<G>(lux_any<G>(code != 0))) {
                // no tile of this chunk acts (the common case: a tile rests ten turns after every action)
                if (valid && cd > 0) M.cells[cell].tile_cd = (uint8_t)(cd - 1);
                continue;
            }
            int cand = valid && (code == LUX_TA_BUILD_WORKER || code == LUX_TA_BUILD_CART) && cd < 1;
            unsigned m0 = lux_ballot<G>(cand && team == 0), m1 = lux_ballot<G>(cand && team == 1);
            int rank = team ? acc1 + lux_popc(m1 & lt) : acc0 + lux_popc(m0 & lt);
            int ok = cand && ((team ? tu1 : tu0) + rank < (team ? tt1 : tt0));
            acc0 += lux_popc(m0); acc1 += lux_popc(m1);
            unsigned mok = lux_ballot<G>(ok);
            int srank = spawned + lux_popc(mok & lt);
            spawned += lux_popc(mok);
            int research = valid && code == LUX_TA_RESEARCH &&
                           (!(flags & LUX_STEP_PREVALIDATE) || cd < 1);   // ResearchAction.is_valid actions.py:412-438
            int acted = research;
            if (ok) {
                if (srank < room) {
                    LuxUnit& N = M.units[n_units0 + srank];
                    N.id = uid0 + 1 + srank;
                    N.x = (uint8_t)(cell % S); N.y = (uint8_t)(cell / S);
                    N.team_type = (uint8_t)(team | (code == LUX_TA_BUILD_CART ? 2 : 0));
                    N.cd_q = 0; N.wood = 0; N.coal = 0; N.uranium = 0; N.pad = 0;
                    M.ust[n_units0 + srank] = 0;
                    M.uact[n_units0 + srank] = 0;
                }
                acted = 1;
            }
            int okroom = ok && srank < room;
            wb0 += lux_popc(lux_ballot<G>(okroom && team == 0 && code == LUX_TA_BUILD_WORKER));
            wb1 += lux_popc(lux_ballot<G>(okroom && team == 1 && code == LUX_TA_BUILD_WORKER));
            cb0 += lux_popc(lux_ballot<G>(okroom && team == 0 && code == LUX_TA_BUILD_CART));
            cb1 += lux_popc(lux_ballot<G>(okroom && team == 1 && code == LUX_TA_BUILD_CART));
            rp0 += lux_popc(lux_ballot<G>(research && team == 0));
            rp1 += lux_popc(lux_ballot<G>(research && team == 1));
            if (valid) {
                int cd0 = cd;
                if (acted) cd = LUX_CITY_ACTION_COOLDOWN;
                if (cd > 0) cd -= 1;
                if (cd != cd0) M.cells[cell].tile_cd = (uint8_t)cd;
            }
        }
        int made = spawned < room ? spawned : room;
        if (made < 0) made = 0;
        n_units = n_units0 + made;
        born0 = wb0 + cb0; born1 = wb1 + cb1;
        if (lane == 0) {
            if (spawned > room) h->status |= LUX_ST_UNIT_OVERFLOW;
            h->unit_id_count = uid0 + made;
            h->workers_built[0] += wb0; h->workers_built[1] += wb1;
            h->carts_built[0] += cb0; h->carts_built[1] += cb1;
            h->research[0] += rp0; h->research[1] += rp1;
        }
    }
    lux_sync<G>();

    LUX_PHASE_SYNC(3);
    // ---- unit turns (Worker.turn unit.py:162-197, Cart.turn :234-269)
    // Moves are conflict-free after pruning and nothing reads another unit's position before the unit pass (a unit
    // holds ONE action: the ordered actions below -- transfer by id, build / pillage on the own cell -- belong to units
    // that do not move), so the moves are applied at the head of the unit pass instead of in a pass of their own.
    // transfer / build city / pillage depend on unit order (team A then B == slot order): one pending
    // action per group and round, the groups of the warp in lockstep
    const int nu_max = lux_wmax<G>(n_units);
#pragma unroll 1
    for (int base = 0; base < nu0_max; base += G) {
        int u = base + lane;
        int k = u < n_units0 ? (M.ust[u] & 7) : 0;
        unsigned pend = lux_ballot<G>(k == LUX_UA_TRANSFER || k == LUX_UA_BCITY || k == LUX_UA_PILLAGE);
#pragma unroll 1
        while (lux_wany<G>(pend != 0)) {
            int su = 0, kind = 0;
            if (pend) {
                su = base + lux_ffs(pend) - 1;
                pend &= pend - 1;
                kind = M.ust[su] & 7;
            }
            if (lux_wany<G>(kind == LUX_UA_TRANSFER)) {             // transfer_resources game.py:1039-1057
                const int xfer = kind == LUX_UA_TRANSFER;
                unsigned w = xfer ? M.uact[su] : 0u;
                int dest_id = (int)(w >> 17);
                int steam = xfer ? unit_team(M.units[su]) : 0;
                int dest = -1;
#pragma unroll 1
                for (int cb = 0; cb < nu_max; cb += G) {
                    int v = cb + lane;
                    int hit = xfer && v < n_units && M.units[v].id == dest_id && unit_team(M.units[v]) == steam;
                    unsigned hm = lux_ballot<G>(hit);
                    if (hm && dest < 0) dest = cb + lux_ffs(hm) - 1;
                }
                if (xfer && lane == 0) {
                    int r = (int)((w >> 3) & 3);
                    if (dest < 0 || r > 2) {
                        M.ust[su] = (uint8_t)(kind | UST_FAILED);
                    } else {
                        LuxUnit& A = M.units[su];
                        LuxUnit& B = M.units[dest];
                        uint16_t* ca = r == 0 ? &A.wood : r == 1 ? &A.coal : &A.uranium;
                        int amt = (int)((w >> 6) & 2047);
                        if (*ca < amt) amt = *ca;
                        int sp = unit_space(B);
                        if (sp < amt) amt = sp;
                        *ca = (uint16_t)(*ca - amt);
                        uint16_t* cbp = r == 0 ? &B.wood : r == 1 ? &B.coal : &B.uranium;
                        *cbp = (uint16_t)(*cbp + amt);
                    }
                }
            }
            if ((kind == LUX_UA_BCITY || kind == LUX_UA_PILLAGE) && lane == 0) {
                LuxUnit& U = M.units[su];
                if (!unit_is_cart(U)) {
                    if (kind == LUX_UA_BCITY) {
                        lux_build_city_lane0<kS, G>(M, su, n_cities, n_tiles, (flags & LUX_STEP_PREVALIDATE) != 0);
                        int spent = 0;                                // expend_resources_for_city unit.py:146-160
                        uint16_t* cg[3] = {&U.wood, &U.coal, &U.uranium};
#pragma unroll 1
                        for (int r = 0; r < 3; r++) {
                            if (spent + *cg[r] > LUX_CITY_BUILD_COST) { *cg[r] = (uint16_t)(*cg[r] - (LUX_CITY_BUILD_COST - spent)); break; }
                            spent += *cg[r];
                            *cg[r] = 0;
                        }
                    } else {                                          // pillage unit.py:188-192
                        int pc = U.y * S + U.x;
                        int rq = LUX_LDG8(&M.cells[pc].road_q);
                        M.cells[pc].road_q = (uint8_t)(rq > LUX_PILLAGE_Q ? rq - LUX_PILLAGE_Q : 0);
                        M.umine[su] = (uint8_t)rq;            // the level it found (read back by the stack path only)
                    }
                }
            }
            lux_sync<G>();
        }
    }
    n_cities = lux_shfl<G>(n_cities, 0);
    n_tiles = lux_shfl<G>(n_tiles, 0);
    if (n_tiles != n_tiles0) tiles_dirty = 1;
    lux_city_prefix<G>(M, n_cities);     // also syncs

    LUX_PHASE_SYNC(4);
    // ---- the unit pass: cooldowns of the acting units, the carts' automatic roads (unit.py:196-197,
    // :257-269), and the per-unit caches the rest of the turn works from -- what the unit stands on
    // (uinfo), which of its five cells it can mine (umin, 2 bits of resource type per direction) and the
    // occupant mark of its cell.  This is the one place a unit reads the grid: five independent loads.
    int stacked = 0;                       // some non-city cell holds several units (rare; see below)
    unsigned minable = 0;                  // bit t: some worker of the match has a cell of resource type t in reach
    {
        int rb0 = 0, rb1 = 0, bad = 0;
        unsigned has_t1 = 0, has_t2 = 0, has_t3 = 0;
#pragma unroll 1
        for (int u = lane; u < n_units; u += G) {
            LuxUnit& U = M.units[u];
            const int isold = u < n_units0;
            unsigned st = isold ? M.ust[u] : 0u;
            if (isold) {
                // the movement flags have served (set in the validation pass, or -- with pre-validation -- only when a
                // move was left): clear the byte of the cell the unit stood on and of the cell it aimed at
                M.p1[U.y * S + U.x] = 0;
                M.p1[M.utgt[u] & UTGT_CELL] = 0;
                if ((st & 7) == LUX_UA_MOVE) {
                    const int dir = (st >> 3) & 7;
                    if ((st & UST_REVERTED) || dir == LUX_DIR_CENTER) st = 0;                      // game.py:462-465
                    else { U.x = (uint8_t)(U.x + lux_dx(dir)); U.y = (uint8_t)(U.y + lux_dy(dir)); }
                }
            }
            const int cart = unit_is_cart(U);
            const int ucell = U.y * S + U.x;
            int pos[5];
            LuxCell nb[5];
            pos[0] = ucell;
            nb[0] = lux_ldcell(M.cells, ucell);
            if (!cart) {
#pragma unroll
                for (int j = 1; j < 5; j++) {
                    pos[j] = lux_cell5(S, U.x, U.y, j);
                    if (pos[j] >= 0) nb[j] = lux_ldcell(M.cells, pos[j]);
                }
            }
            if ((st & 7) != 0 && !(st & UST_FAILED)) U.cd_q = (uint8_t)(U.cd_q + (cart ? 12 : 8) * mult);
            const LuxCell& C = nb[0];
            unsigned info = 0xFFFFu;
            if (cell_is_tile(C)) {
                info = (unsigned)(M.cprefix[C.city_slot] + C.key) | ((unsigned)C.city_slot << 16) | UI_TILE |
                       (cell_tile_team(C) == unit_team(U) ? UI_OWN : 0u);
            } else {
                if (cart) {
                    // the road level this cart found (and whether its turn was cut short): should the cell turn out
                    // to hold several carts, lane 0 redoes their road steps one after the other (below)
                    M.umine[u] = (uint8_t)(C.road_q | ((st & UST_FAILED) ? 0x80u : 0u));
                    if (!(st & UST_FAILED) && C.road_q < LUX_MAX_ROAD_Q) {
                        int r = C.road_q + LUX_ROAD_DEV_Q;
                        M.cells[ucell].road_q = (uint8_t)(r > LUX_MAX_ROAD_Q ? LUX_MAX_ROAD_Q : r);
                        if (unit_team(U)) rb1 += LUX_ROAD_DEV_Q; else rb0 += LUX_ROAD_DEV_Q;
                    }
                }
                // a non-city cell holds one unit in play; a stack (left behind by a destroyed city) is rare and
                // gets its occupant chain rebuilt below
                if (lux_atomic_or_byte(M.occ, ucell, (unsigned)(u + 1))) bad = 1;
            }
            unsigned mm = 0;
            if (!cart) {
#pragma unroll
                for (int j = 0; j < 5; j++)
                    if (pos[j] >= 0 && nb[j].amount > 0) mm |= (unsigned)cell_res_type(nb[j]) << (2 * j);
            }
            M.uinfo[u] = info;
            M.umin[u] = (uint16_t)mm;
            M.ust[u] = (uint8_t)(st & 7);          // flags off (the night marks the dead with UST_FAILED); the kind stays for the stack path
            {   // which resource types occur among the five fields (type = hi:lo of a field)
                const unsigned lo = mm & 0x155u, hi = (mm >> 1) & 0x155u;
                has_t1 |= lo & ~hi; has_t2 |= hi & ~lo; has_t3 |= lo & hi;
            }
        }
        minable = (lux_any<G>(has_t1 != 0) ? 1u << LUX_RES_WOOD : 0u) | (lux_any<G>(has_t2 != 0) ? 1u << LUX_RES_COAL : 0u) |
                  (lux_any<G>(has_t3 != 0) ? 1u << LUX_RES_URANIUM : 0u);
        {   // one reduction: road steps are < 2^12 per team (3 per cart), stacks < 2^8
            int packed = lux_reduce_add<G>(rb0 | (rb1 << 12) | (bad << 24));
            rb0 = packed & 0xFFF; rb1 = (packed >> 12) & 0xFFF; bad = packed >> 24;
        }
        if (lane == 0) h->roads_built_q[0] += rb0, h->roads_built_q[1] += rb1;
        stacked = bad;
        if (lux_wany<G>(bad)) {
            // several units share a non-city cell (game.py:1173: a city fell under them and they survived the
            // night): the OR-ed occupant bytes of those cells are meaningless; lane 0 rebuilds them as chains
            // occ[cell] -> unit -> unext -> ... in unit order
            lux_sync<G>();
            if (bad && lane == 0) lux_stack_paths<kS>(M, n_units);
        }
    }
    lux_sync<G>();

    LUX_PHASE_SYNC(5);
    // ---- resource collection: uranium, coal, wood (distribute_all_resources game.py:868-880)
    {
        const int rp0 = h->research[0], rp1 = h->research[1];
        // the passes worth running: a team has researched the type AND some worker stands next to a cell of it
        // (otherwise the pass finds no request: every umine = none, nothing marked, nothing gained)
        unsigned run = minable & ((1u << LUX_RES_WOOD) | ((rp0 >= LUX_RESEARCH_COAL || rp1 >= LUX_RESEARCH_COAL) ? 1u << LUX_RES_COAL : 0u) |
                                  ((rp0 >= LUX_RESEARCH_URANIUM || rp1 >= LUX_RESEARCH_URANIUM) ? 1u << LUX_RES_URANIUM : 0u));
        if (G != 32)     // the groups of a warp walk through the union of their passes
            run = (lux_any_warp(run & 2u) ? 2u : 0u) | (lux_any_warp(run & 4u) ? 4u : 0u) | (lux_any_warp(run & 8u) ? 8u : 0u);
#pragma unroll 1
        while (run) {
            const int rtype = (run & 8u) ? LUX_RES_URANIUM : (run & 4u) ? LUX_RES_COAL : LUX_RES_WOOD;
            run &= ~(1u << rtype);
            const int need = rtype == LUX_RES_WOOD ? 0 : rtype == LUX_RES_COAL ? LUX_RESEARCH_COAL : LUX_RESEARCH_URANIUM;
            lux_collect_type<kS, G>(M, rtype, n_units, rp0 >= need, rp1 >= need, stacked);
        }
    }

    LUX_PHASE_SYNC(6);
    // ---- deposit (handle_resource_deposit game.py:1003-1023)
    {
        int fg0 = 0, fg1 = 0;
#pragma unroll 1
        for (int u = lane; u < n_units; u += G) {
            LuxUnit& U = M.units[u];
            unsigned info = M.uinfo[u];
            M.occ[U.y * S + U.x] = 0;               // collection is over: drop the occupant mark again
            if (info & UI_OWN) {
                int fuel = U.wood + 10 * U.coal + 40 * U.uranium;
                if (fuel) lux_atomic_add(&M.cities[(info >> 16) & 0xFFu].fuel, fuel);
                if (unit_team(U)) fg1 += fuel; else fg0 += fuel;
                U.wood = 0; U.coal = 0; U.uranium = 0;
            }
        }
        fg0 = lux_reduce_add<G>(fg0);
        fg1 = lux_reduce_add<G>(fg1);
        if (lane == 0) { h->fuel_generated[0] += fg0; h->fuel_generated[1] += fg1; }
    }
    lux_sync<G>();

    LUX_PHASE_SYNC(7);
    // ---- night (handle_night game.py:537-558); matches of one warp differ in their turn counters, so
    // the phase runs when any of them has night and the others walk through it with empty lists
    if (lux_wany<G>(night)) {
        const int nc_n = night ? n_cities : 0, nu_n = night ? n_units : 0;
        int died = 0;
#pragma unroll 1
        for (int s = lane; s < nc_n; s += G) {
            LuxCity& city = M.cities[s];
            M.cremap[s] = 0;
            if (city.ntiles == 0) continue;
            int upkeep = LUX_UPKEEP_CITY * city.ntiles - LUX_ADJ_BONUS * city.adjsum;   // city.py:34-50
            if (city.fuel < upkeep) { M.cremap[s] = 1; died = 1; }
            else city.fuel -= upkeep;
        }
        lux_sync<G>();
        died = lux_any<G>(died);
        if (died) {
            tiles_dirty = 1;
#pragma unroll 1
            for (int t = lane; t < n_tiles; t += G) {                // destroy_city :1059-1070
                int tc = M.tiles[t];
                LuxCell C = lux_ldcell(M.cells, tc);
                if (cell_is_tile(C) && M.cremap[C.city_slot]) {
                    C.flags &= 0x03;
                    C.tile_cd = 0; C.city_slot = 0; C.key = 0;
                    C.road_q = 0;
                    lux_stcell(M.cells, tc, C);
                }
            }
        }
#pragma unroll 1
        for (int u = lane; u < nu_n; u += G) {                        // spend_fuel_to_survive unit.py:64-98
            LuxUnit& U = M.units[u];
            unsigned info = M.uinfo[u];
            if (info & UI_TILE) {
                if (!died || !M.cremap[(info >> 16) & 0xFFu]) continue;
                M.uinfo[u] = 0xFFFFu;                                  // its city fell: it stands in the open now
            }
            int need = unit_is_cart(U) ? LUX_UPKEEP_CART : LUX_UPKEEP_WORKER;
            int used = U.wood < need ? U.wood : need;
            need -= used; U.wood = (uint16_t)(U.wood - used);
            if (need > 0) {
                int want = (need + 9) / 10;
                used = U.coal < want ? U.coal : want;
                need -= used * 10; U.coal = (uint16_t)(U.coal - used);
            }
            if (need > 0) {
                int want = (need + 39) / 40;
                used = U.uranium < want ? U.uranium : want;
                need -= used * 40; U.uranium = (uint16_t)(U.uranium - used);
            }
            if (need > 0) {                                           // dead: dropped while storing
                M.ust[u] = UST_FAILED;
                if (unit_team(U)) died1++; else died0++;
            }
        }
        died0 = lux_reduce_add<G>(died0);
        died1 = lux_reduce_add<G>(died1);
        lux_sync<G>();
        if (died) {
#pragma unroll 1
            for (int s = lane; s < n_cities; s += G)
                if (M.cremap[s]) { M.cities[s].ntiles = 0; M.cities[s].fuel = 0; M.cities[s].adjsum = 0; }
        }
        lux_sync<G>();
    }

    LUX_PHASE_SYNC(8);
    // ---- depleted cells leave the list (game.py:498-510), trees regrow (:1081-1102).  The resource list
    // carries each cell's type and a mirror of its amount, so this pass over ALL resource cells works on shared
    // memory; only a cell somebody mined from this turn is re-read from the grid (the collection pass stored
    // its new amount there), and only changed amounts are written to the grid.
    {
        const int n_res = h->n_res;
#pragma unroll 1
        for (int r = lane; r < n_res; r += G) {
            const LuxRes e = M.res[r];
            const int cell = LUX_RES_CELL(e);
            const int before = LUX_RES_AMOUNT(e);
            const unsigned mark = M.p1[cell];
            int a = before;
            if (mark) {
                M.p1[cell] = 0;                                              // collection scratch (mark + fill level)
                if (mark & 0x80u) a = LUX_LDG16(&M.cells[cell].amount);
            }
            if (a == 0) {
                // depleted this turn: the cell loses its resource type (amount + flags share one 32-bit word)
                if (before > 0) lux_atomic_and((unsigned*)&M.cells[cell], ~0x00030000u);
            } else if (LUX_RES_TYPE(e) == LUX_RES_WOOD && a < LUX_MAX_WOOD) {
                a = (41 * a + 39) / 40;                                      // == ceil(min(a*1.025, 500)), tests/test_oracle_golden.py
                if (a > LUX_MAX_WOOD) a = LUX_MAX_WOOD;
                M.cells[cell].amount = (uint16_t)a;
            }
            // nothing reads the staged list after this pass: a changed entry goes straight to the list in global memory
            // (the unchanged ones -- coal, uranium, full or unmined wood -- are not written at all)
            if (a != before) M.g_res[r] = (e & 0xFFFFu) | ((LuxRes)a << 16);
        }
    }
    lux_sync<G>();

    LuxTurnOut out;
    out.n_units = n_units; out.n_cities = n_cities; out.n_tiles = n_tiles; out.tiles_dirty = tiles_dirty;
    out.alive0 = tu0 + born0 - died0; out.alive1 = tu1 + born1 - died1;
    out.tt0 = tt0; out.tt1 = tt1;
    return out;
}

// ---- epilogue: compaction, match_over, cooldowns, canonical tile order, store ---------------
// g_* are the match's sections in global memory (the grid is already up to date there).
template <int kS, int G>
LUX_DEV void lux_finish_and_store(LuxMatch& M, const LuxTurnOut& T, LuxHeader* g_hdr, LuxUnit* g_units,
                                  LuxCity* g_cities, uint16_t* g_tiles, LuxRes* g_res, int active) {
    const int lane = lux_lane<G>();
    const unsigned lt = lux_lanemask_lt<G>();
    LuxHeader* h = M.hdr;
    const int S = kS > 0 ? kS : M.S;

    // (the resource list: the regrowth pass has written its changed entries in place)
    (void)g_res;
    // Cities.  The common turn neither founds, merges nor loses a city tile (tiles_dirty == 0: the tile list did not
    // grow and no city fell): the cities go back as they stand and the per-team tile counts are the turn's opening
    // ones.  Otherwise: drop the dead (ntiles == 0), keep order, rebuild the canonical tile list.
    int n_alive = T.n_cities, tl0 = T.tt0, tl1 = T.tt1;
    if (!lux_wany<G>(T.tiles_dirty)) {
        if (active) lux_copy16<G>(g_cities, M.cities, T.n_cities * 16);
    } else {
    n_alive = 0; tl0 = 0; tl1 = 0;
    const int nc_max = lux_wmax<G>(T.n_cities);
#pragma unroll 1
    for (int base = 0; base < nc_max; base += G) {
        int s = base + lane;
        int alive = s < T.n_cities && M.cities[s].ntiles > 0;
        unsigned m = lux_ballot<G>(alive);
        if (s < T.n_cities) M.cremap[s] = (uint8_t)(alive ? n_alive + lux_popc(m & lt) : 0xFF);
        n_alive += lux_popc(m);
        if (alive) { if (M.cities[s].team) tl1 += M.cities[s].ntiles; else tl0 += M.cities[s].ntiles; }
    }
    {   // one reduction (tiles <= 1024 per team)
        unsigned pt = (unsigned)lux_reduce_add<G>(tl0 | (tl1 << 16));
        tl0 = (int)(pt & 0xFFFFu); tl1 = (int)(pt >> 16);
    }
    lux_sync<G>();
    const int cities_moved = n_alive != T.n_cities;
    // store cities compacted
#pragma unroll 1
    for (int s = lane; s < T.n_cities; s += G) {
        int ns = M.cremap[s];
        if (ns != 0xFF) g_cities[ns] = M.cities[s];
    }
    // canonical tile list: position = prefix[new slot] + key; surviving tiles of moved cities get their
    // new slot written into the grid in the same pass
    const int relist = T.tiles_dirty || cities_moved;
    if (lux_wany<G>(relist)) {
        const int nt_r = relist ? T.n_tiles : 0;
        int run = 0;
#pragma unroll 1
        for (int base = 0; base < nc_max; base += G) {
            int s = base + lane;
            int v = (s < T.n_cities) ? M.cities[s].ntiles : 0;   // dead cities have 0 tiles
            int tot;
            int ex = lux_exscan_add<G>(v, &tot);
            if (relist && s < T.n_cities && v > 0) M.cprefix[M.cremap[s]] = (uint16_t)(run + ex);
            run += tot;
        }
        lux_sync<G>();
#pragma unroll 1
        for (int t = lane; t < nt_r; t += G) {
            int cell = M.tiles[t];
            LuxCell C = lux_ldcell(M.cells, cell);
            if (cell_is_tile(C)) {
                int ns = M.cremap[C.city_slot];
                if (ns != C.city_slot) M.cells[cell].city_slot = (uint8_t)ns;
                g_tiles[M.cprefix[ns] + C.key] = (uint16_t)cell;
            }
        }
    }
    }
    const int n_tiles_live = tl0 + tl1;

    // match_over (game.py:570-592) is evaluated before the turn counter moves (:515-517)
    const int alive0 = T.alive0, alive1 = T.alive1;    // start-of-turn counts + spawns - night deaths
    // (a team holds a city exactly when it holds a city tile)
    int over = h->turn >= LUX_MAX_DAYS - 1 || (alive0 + tl0 == 0) || (alive1 + tl1 == 0);

    // run_cooldowns (game.py:560-568) and the stable partition team A | team B while storing
    int run0 = 0, run1 = 0;
    const int nu_max = lux_wmax<G>(T.n_units);
#pragma unroll 1
    for (int base = 0; base < nu_max; base += G) {
        int u = base + lane;
        int valid = u < T.n_units && !(M.ust[u] & UST_FAILED);
        int team = valid ? unit_team(M.units[u]) : 0;
        unsigned m0 = lux_ballot<G>(valid && team == 0), m1 = lux_ballot<G>(valid && team == 1);
        if (valid) {
            LuxUnit U = M.units[u];
            int road = (M.uinfo[u] & UI_TILE) ? LUX_MAX_ROAD_Q : (int)LUX_LDG8(&M.cells[U.y * S + U.x].road_q);    // cell.py:77-85
            int cd = (int)U.cd_q - road - 4;
            U.cd_q = (uint8_t)(cd > 0 ? cd : 0);
            U.pad = 0;
            int dst = team ? alive0 + run1 + lux_popc(m1 & lt) : run0 + lux_popc(m0 & lt);
            g_units[dst] = U;
        }
        run0 += lux_popc(m0); run1 += lux_popc(m1);
    }

    if (lane == 0) {
        h->turn += 1;
        h->n_units = (uint16_t)(alive0 + alive1);
        h->n_team_units[0] = (uint16_t)alive0;
        h->n_team_units[1] = (uint16_t)alive1;
        h->n_cities = (uint16_t)n_alive;
        h->n_tiles = (uint16_t)n_tiles_live;
        h->n_team_tiles[0] = (uint16_t)tl0;
        h->n_team_tiles[1] = (uint16_t)tl1;
        if (over) {
            h->done = 1;
            int w;                                                     // get_winning_team game.py:594-634
            if (tl0 != tl1) w = tl0 > tl1 ? LUX_WIN_A : LUX_WIN_B;
            else if (alive0 != alive1) w = alive0 > alive1 ? LUX_WIN_A : LUX_WIN_B;
            else if (h->fuel_generated[0] != h->fuel_generated[1])
                w = h->fuel_generated[0] > h->fuel_generated[1] ? LUX_WIN_A : LUX_WIN_B;
            else w = LUX_WIN_TIE;
            h->winner = (uint8_t)w;
        }
    }
    lux_sync<G>();
    // header: a straight copy
    if (active) lux_copy16<G>(g_hdr, M.hdr, LUX_HDR_BYTES);
}
