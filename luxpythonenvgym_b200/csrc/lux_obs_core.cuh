// lux_obs_core.cuh -- batched per-entity observation vectors, one warp per match.
//
// Replaces AgentPolicy.get_observation (reference: examples/agent_policy.py:188-424), which the
// reference calls once per actionable entity from LuxEnvironment.step (luxai2021/env/lux_env.py:146).
// Here one warp produces the 85-float rows of every entity MatchController would yield for the
// selected team (match_controller.py:271-289) in one pass:
//   * the per-turn node lists (agent_policy.py:197-247) are built once per match in shared memory by
//     ballot compaction (wood / coal / uranium cells in GameMap.resources order, own city tiles in
//     cities x cells order, own workers in dict order);
//   * each lane owns one entity and scans the lists once per key, keeping the lexicographic
//     (distance, index) minimum, second minimum, maximum and second maximum -- that gives both the
//     closest and the furthest node with numpy's first-index tie-break, with and without the
//     np.delete(closest) self-exclusion (:328-336);
//   * rows are assembled in a shared-memory staging tile of 32 entities and written with 16-byte
//     coalesced stores.
// Values are computed in float64 exactly as the reference's Python expressions and rounded once.
#pragma once
#include "lux_step_core.cuh"

#define LUX_OBS_DIM 85

struct LuxObsLayout {
    int hdr, cells, units, cities, tiles, res, nodes_res, nodes_tile, nodes_work, first, count, stage, total;
};

#ifdef LUX_EMULATE
static inline
#else
__host__ __device__ inline
#endif
LuxObsLayout lux_obs_layout(int size, int u_cap, int c_cap, int t_cap, int r_cap) {
    LuxObsLayout L;
    int ncell = size * size;
    int o = 0;
#define LUX_TAKE(field, bytes) L.field = o; o += (((bytes) + 15) & ~15);
    LUX_TAKE(hdr, LUX_HDR_BYTES)
    LUX_TAKE(cells, ncell * 8)
    LUX_TAKE(units, u_cap * 16)
    LUX_TAKE(cities, c_cap * 16)
    LUX_TAKE(tiles, t_cap * 2)
    LUX_TAKE(res, r_cap * 2)
    LUX_TAKE(nodes_res, r_cap * 2)      // wood | coal | uranium node positions (x | y << 8)
    LUX_TAKE(nodes_tile, t_cap * 2)
    LUX_TAKE(nodes_work, u_cap * 2)
    LUX_TAKE(first, ncell * 4)          // lowest unit slot standing on the cell (0xFFFFFFFF none)
    LUX_TAKE(count, ncell)              // units standing on the cell
    LUX_TAKE(stage, 32 * LUX_OBS_DIM * 4)
#undef LUX_TAKE
    L.total = o;
    return L;
}

struct LuxObsMatch {
    LuxHeader* hdr; LuxCell* cells; LuxUnit* units; LuxCity* cities; uint16_t* tiles; uint16_t* res;
    uint16_t* nodes_res; uint16_t* nodes_tile; uint16_t* nodes_work;
    unsigned* first; uint8_t* count; float* stage;
    int S, ncell;
};

LUX_DEV void lux_obs_bind(LuxObsMatch& M, unsigned char* smem, const LuxObsLayout& L, int size) {
    M.hdr = (LuxHeader*)(smem + L.hdr); M.cells = (LuxCell*)(smem + L.cells);
    M.units = (LuxUnit*)(smem + L.units); M.cities = (LuxCity*)(smem + L.cities);
    M.tiles = (uint16_t*)(smem + L.tiles); M.res = (uint16_t*)(smem + L.res);
    M.nodes_res = (uint16_t*)(smem + L.nodes_res); M.nodes_tile = (uint16_t*)(smem + L.nodes_tile);
    M.nodes_work = (uint16_t*)(smem + L.nodes_work);
    M.first = (unsigned*)(smem + L.first); M.count = (uint8_t*)(smem + L.count);
    M.stage = (float*)(smem + L.stage);
    M.S = size; M.ncell = size * size;
}

// lexicographic trackers over (squared distance, list index)
struct LuxArg { int d, i; };
LUX_DEV void lux_arg_min2(LuxArg& a1, LuxArg& a2, int d, int i) {
    if (d < a1.d) { a2 = a1; a1.d = d; a1.i = i; }
    else if (d < a2.d) { a2.d = d; a2.i = i; }
}
LUX_DEV void lux_arg_max2(LuxArg& a1, LuxArg& a2, int d, int i) {
    if (d > a1.d) { a2 = a1; a1.d = d; a1.i = i; }
    else if (d > a2.d) { a2.d = d; a2.i = i; }
}

LUX_DEV unsigned lux_atomic_min_u32(unsigned* p, unsigned v) {
#ifdef LUX_EMULATE
    unsigned o = *p; if (v < o) *p = v; return o;
#else
    return atomicMin(p, v);
#endif
}

// one (closest|furthest, key) block of 7 floats (agent_policy.py:327-385)
LUX_DEV void lux_obs_block(const LuxObsMatch& M, float* o, const uint16_t* nodes, int n, int px, int py,
                           int exclude, int furthest, int key) {
    if (n == 0) return;                                   // key not in object_nodes
    LuxArg mn1 = {0x7fffffff, -1}, mn2 = {0x7fffffff, -1}, mx1 = {-1, -1}, mx2 = {-1, -1};
    for (int i = 0; i < n; i++) {
        int xy = nodes[i];
        int dx = (xy & 0xff) - px, dy = (xy >> 8) - py;
        int d = dx * dx + dy * dy;
        lux_arg_min2(mn1, mn2, d, i);
        lux_arg_max2(mx1, mx2, d, i);
    }
    int pick;
    if (exclude) {
        if (n == 1) { o[5] = 1.0f; return; }              // nothing left after np.delete
        if (!furthest) pick = mn2.i;
        else pick = (mx1.i == mn1.i) ? mx2.i : mx1.i;
    } else {
        pick = furthest ? mx1.i : mn1.i;
    }
    int xy = nodes[pick];
    int tx = xy & 0xff, ty = xy >> 8;
    // Position.direction_to (position.py:48-66): N, E, S, W, first strict improvement; slots c,n,w,s,e
    int slot = 0;
    if (ty < py) slot = 1; else if (tx > px) slot = 4; else if (ty > py) slot = 3; else if (tx < px) slot = 2;
    o[slot] = 1.0f;
    int man = (tx > px ? tx - px : px - tx) + (ty > py ? ty - py : py - ty);
    double dist = (double)man / 20.0;
    o[5] = (float)(dist < 1.0 ? dist : 1.0);
    const LuxCell& T = M.cells[ty * M.S + tx];
    double v;
    if (key == 3) {
        const LuxCity& city = M.cities[T.city_slot];
        double upkeep = (double)((int)city.ntiles * LUX_UPKEEP_CITY - (int)city.adjsum * LUX_ADJ_BONUS);
        v = (double)city.fuel / (upkeep * 200.0);
    } else if (key < 3) {
        v = (double)T.amount / 500.0;
    } else {
        const LuxUnit& F = M.units[M.first[ty * M.S + tx]];
        v = (double)unit_space(F) / 100.0;
    }
    o[6] = (float)(v < 1.0 ? v : 1.0);
}

// Shared memory holds the match (hdr, cells, units, cities, tiles, res).  Writes obs rows, entity
// codes and the entity count for `team`.  Returns the number of entities (clamped to e_cap).
LUX_DEV int lux_observe(LuxObsMatch& M, int team, float* g_obs, int16_t* g_ent, int e_cap) {
    const int lane = lux_lane();
    const unsigned lt = lux_lanemask_lt();
    const LuxHeader* h = M.hdr;
    const int S = M.S;
    const int n_units = h->n_units, n_tiles = h->n_tiles, n_res = h->n_res;

    // ---- per-cell unit count and lowest slot
    for (int i = lane; i < M.ncell; i += 32) M.first[i] = 0xFFFFFFFFu;
    lux_zero16(M.count, lux_align16(M.ncell));
    lux_sync();
    for (int u = lane; u < n_units; u += 32) {
        int c = M.units[u].y * S + M.units[u].x;
        lux_atomic_min_u32(&M.first[c], (unsigned)u);
        unsigned* w = (unsigned*)(M.count + (c & ~3));
        lux_atomic_add((int*)w, 1 << ((c & 3) * 8));
    }

    // ---- node lists (agent_policy.py:197-247), order-preserving ballot compaction
    int n_type[3] = {0, 0, 0};
    {
        int base_out = 0;
        for (int t = 1; t <= 3; t++) {
            int cnt = 0;
            for (int base = 0; base < n_res; base += 32) {
                int r = base + lane;
                int cell = r < n_res ? M.res[r] : 0;
                int ok = r < n_res && M.cells[cell].amount > 0 && cell_res_type(M.cells[cell]) == t;
                unsigned m = lux_ballot(ok);
                if (ok) M.nodes_res[base_out + cnt + lux_popc(m & lt)] = (uint16_t)((cell % S) | ((cell / S) << 8));
                cnt += lux_popc(m);
            }
            n_type[t - 1] = cnt;
            base_out += cnt;
        }
    }
    int own_tiles = 0, opp_tiles = 0;
    for (int base = 0; base < n_tiles; base += 32) {
        int p = base + lane;
        int cell = p < n_tiles ? M.tiles[p] : 0;
        int own = p < n_tiles && cell_tile_team(M.cells[cell]) == team;
        unsigned m = lux_ballot(own);
        if (own) M.nodes_tile[own_tiles + lux_popc(m & lt)] = (uint16_t)((cell % S) | ((cell / S) << 8));
        own_tiles += lux_popc(m);
        opp_tiles += lux_popc(lux_ballot(p < n_tiles && !own));
    }
    int own_workers = 0, own_carts = 0;
    for (int base = 0; base < n_units; base += 32) {
        int u = base + lane;
        int mine = u < n_units && unit_team(M.units[u]) == team;
        int worker = mine && !unit_is_cart(M.units[u]);
        unsigned m = lux_ballot(worker);
        if (worker) M.nodes_work[own_workers + lux_popc(m & lt)] = (uint16_t)(M.units[u].x | (M.units[u].y << 8));
        own_workers += lux_popc(m);
        own_carts += lux_popc(lux_ballot(mine && !worker));
    }
    lux_sync();

    const int night = (h->turn % LUX_CYCLE_LENGTH) >= LUX_DAY_LENGTH;
    const float f_turn = (float)((double)h->turn / 360.0);
    const float f_own_tiles = (float)((double)own_tiles / 30.0), f_opp_tiles = (float)((double)opp_tiles / 30.0);
    const float f_workers = (float)((double)own_workers / 30.0), f_carts = (float)((double)own_carts / 30.0);
    const int rp = h->research[team];
    const float f_rp = (float)((double)rp / 200.0);
    const uint16_t* nodes_key[5] = {M.nodes_res, M.nodes_res + n_type[0], M.nodes_res + n_type[0] + n_type[1],
                                    M.nodes_tile, M.nodes_work};
    const int n_key[5] = {n_type[0], n_type[1], n_type[2], own_tiles, own_workers};

    // ---- entities: the team's units with can_act, then its tiles with can_act, 32 per pass
    int n_ent = 0;
    const int total = n_units + n_tiles;
    for (int base = 0; base < total; base += 32) {
        int e = base + lane;
        int is_unit = e < n_units;
        int px = 0, py = 0, kind = 0, code = 0, valid = 0, alone = 0;
        if (is_unit) {
            const LuxUnit& U = M.units[e];
            valid = unit_team(U) == team && U.cd_q < 4;
            px = U.x; py = U.y; kind = unit_is_cart(U) ? 1 : 0; code = e;
            alone = M.count[py * S + px] <= 1;
        } else if (e < total) {
            int p = e - n_units;
            int cell = M.tiles[p];
            const LuxCell& C = M.cells[cell];
            valid = cell_tile_team(C) == team && C.tile_cd < 1;
            px = cell % S; py = cell / S; kind = 2; code = 0x4000 | p;
        }
        unsigned vm = lux_ballot(valid);
        int nv = lux_popc(vm);
        if (nv == 0) continue;
        int row = lux_popc(vm & lt);
        // zero the staging rows that will be used
        for (int i = lane; i < nv * LUX_OBS_DIM; i += 32) M.stage[i] = 0.0f;
        lux_sync();
        if (valid) {
            float* o = M.stage + row * LUX_OBS_DIM;
            o[kind] = 1.0f;
            int at = 3;
            for (int furthest = 0; furthest < 2; furthest++)
                for (int key = 0; key < 5; key++, at += 7) {
                    int exclude = (key == 3 && kind == 2) || (key == 4 && kind == 0 && alone);
                    lux_obs_block(M, o + at, nodes_key[key], n_key[key], px, py, exclude, furthest, key);
                }
            if (kind != 2) o[73] = (float)((double)unit_space(M.units[e]) / 100.0);
            o[74] = night ? 1.0f : 0.0f;
            o[75] = f_turn;
            o[76] = f_own_tiles; o[77] = f_opp_tiles;
            o[78] = f_workers; o[79] = f_workers;       // agent_policy.py:214-215: the "opponent" lists hold the own units
            o[80] = f_carts; o[81] = f_carts;
            o[82] = f_rp;
            o[83] = rp >= LUX_RESEARCH_COAL ? 1.0f : 0.0f;
            o[84] = rp >= LUX_RESEARCH_URANIUM ? 1.0f : 0.0f;
            if (n_ent + row < e_cap) g_ent[n_ent + row] = (int16_t)code;
        }
        lux_sync();
        // coalesced copy of nv rows to obs[n_ent .. n_ent+nv)
        int room = e_cap - n_ent;
        int rows_out = nv < room ? nv : (room > 0 ? room : 0);
        float* dst = g_obs + (size_t)n_ent * LUX_OBS_DIM;
        for (int i = lane; i < rows_out * LUX_OBS_DIM; i += 32) dst[i] = M.stage[i];
        n_ent += nv;
        lux_sync();
    }
    return n_ent;
}
