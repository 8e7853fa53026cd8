// lux_obs_core.cuh -- batched per-entity observation vectors, one warp per match.
//
// Replaces AgentPolicy.get_observation (reference: examples/agent_policy.py:188-424), which the
// reference calls once per actionable entity from LuxEnvironment.step (luxai2021/env/lux_env.py:146).
// Here one warp produces the 85-float rows of every entity MatchController would yield for the
// selected team (match_controller.py:271-289) in one pass:
//   * the per-turn node lists (agent_policy.py:197-247) are built once per match in shared memory by
//     ballot compaction (wood / coal / uranium cells in GameMap.resources order, own city tiles in
//     cities x cells order, own workers in dict order);
//   * the warp takes the actionable entities one at a time (a typical match has one or two) and scans
//     each node list cooperatively, keeping the lexicographic (distance, index) minimum, second minimum,
//     maximum and second maximum -- that gives both the closest and the furthest node with numpy's
//     first-index tie-break, with and without the np.delete(closest) self-exclusion (:328-336); ten lanes
//     then write the ten feature blocks of the row in parallel;
//   * a row is ~70 % zeros: the warp first zero-fills the rows of the 32 entities in flight with
//     coalesced 16-byte stores, then every lane drops its ~25 non-zero features into its own row.
//     No staging tile is needed.
//   * the cell grid is read in place (like the step kernel): one global load per resource cell and per
//     city tile while the node lists are built; what the feature blocks need from a cell (amount, city
//     slot, tile team / cooldown) rides in the node entries, so nothing else touches the grid.
// Values are computed in float64 exactly as the reference's Python expressions and rounded once.
#pragma once
#include "lux_step_core.cuh"

#define LUX_OBS_DIM 85

struct LuxObsLayout {
    int hdr, units, tiles, res, nodes_res, nodes_tile, nodes_work, tinfo, r_cap, total;
};

#ifdef LUX_EMULATE
static inline
#else
__host__ __device__ inline
#endif
LuxObsLayout lux_obs_layout(int size, int u_cap, int t_cap, int r_cap) {
    LuxObsLayout L;
    int ncell = size * size;
    int o = 0;
#define LUX_TAKE(field, bytes) L.field = o; o += (((bytes) + 15) & ~15);
    (void)ncell;
    LUX_TAKE(hdr, LUX_HDR_BYTES)
    LUX_TAKE(units, u_cap * 16)
    LUX_TAKE(tiles, t_cap * 2)
    LUX_TAKE(res, r_cap * 4)
    LUX_TAKE(nodes_res, 3 * r_cap * 4)  // wood, coal, uranium nodes: x | y << 8 | amount << 16, r_cap entries each
    LUX_TAKE(nodes_tile, t_cap * 4)     // own city tiles: x | y << 8 | city slot << 16
    LUX_TAKE(nodes_work, u_cap * 4)     // own workers: x | y << 8
    LUX_TAKE(tinfo, t_cap)              // per tile: team | can_act << 1
#undef LUX_TAKE
    L.r_cap = r_cap;
    L.total = o;
    return L;
}

struct LuxObsMatch {
    LuxHeader* hdr; LuxUnit* units; uint16_t* tiles; LuxRes* res;
    uint32_t* nodes_res; uint32_t* nodes_tile; uint32_t* nodes_work; uint8_t* tinfo;
    const LuxCell* cells;        // the grid stays in global memory, read once per resource cell / city tile
    const LuxCity* g_cities;     // cities stay in global memory: two lookups per entity
    int S, ncell, r_cap;
};

LUX_DEV void lux_obs_bind(LuxObsMatch& M, unsigned char* smem, const LuxObsLayout& L, int size, const LuxCell* g_cells,
                          const LuxCity* g_cities) {
    M.hdr = (LuxHeader*)(smem + L.hdr);
    M.units = (LuxUnit*)(smem + L.units);
    M.tiles = (uint16_t*)(smem + L.tiles); M.res = (LuxRes*)(smem + L.res);
    M.nodes_res = (uint32_t*)(smem + L.nodes_res); M.nodes_tile = (uint32_t*)(smem + L.nodes_tile);
    M.nodes_work = (uint32_t*)(smem + L.nodes_work);
    M.tinfo = smem + L.tinfo;
    M.cells = g_cells;
    M.g_cities = g_cities;
    M.S = size; M.ncell = size * size; M.r_cap = L.r_cap;
}

// float32(a / c) for a small integer a (|a| < 2^16) and an integer constant c <= 500, as the Python expression
// computes it (a double division, rounded once more when the row is stored as float32) -- without the division:
// a * RN64(1 / c) is within 2^-52 of a / c, and a / c is never closer than 2^-25 / c >= 2^-34 (relative) to a
// float32 rounding boundary (a . 2^(24-e) - (2m+1) c is a non-zero integer: a / c cannot carry a 25-bit odd
// significand unless a >= 2^15 c), so both round to the same float32.  A double division costs ~30 instructions on
// the FP64 pipe; the rows need seven of them per entity.  min(v, 1) is applied in float32: RN32 is monotone.
LUX_DEV float lux_ratio32(int a, double inv_c) { return (float)((double)a * inv_c); }
LUX_DEV float lux_ratio32_cap1(int a, double inv_c) {
    const float v = lux_ratio32(a, inv_c);
    return v < 1.0f ? v : 1.0f;
}

// One (closest | furthest, key) block of 7 floats (agent_policy.py:327-385) for the chosen node, written into the
// (zeroed) row by ONE lane.
LUX_DEV void lux_obs_emit(const LuxObsMatch& M, float* o, unsigned node, int px, int py, int key, int n_units) {
    int tx = node & 0xff, ty = (node >> 8) & 0xff;
    // Position.direction_to (position.py:48-66): N, E, S, W, first strict improvement; slots c,n,w,s,e
    int slot = 0;
    if (ty < py) slot = 1; else if (tx > px) slot = 4; else if (ty > py) slot = 3; else if (tx < px) slot = 2;
    o[slot] = 1.0f;
    int man = (tx > px ? tx - px : px - tx) + (ty > py ? ty - py : py - ty);
    o[5] = lux_ratio32_cap1(man, 1.0 / 20.0);
    if (key < 3) {
        o[6] = lux_ratio32_cap1((int)(node >> 16), 1.0 / 500.0);
        return;
    }
    double v;
    if (key == 3) {
        const LuxCity city = M.g_cities[node >> 16];
        double upkeep = (double)((int)city.ntiles * LUX_UPKEEP_CITY - (int)city.adjsum * LUX_ADJ_BONUS);
        v = (double)city.fuel / (upkeep * 200.0);
    } else {
        // first unit in that cell's dict (:380-381): every unit standing on a cell gives the same value in
        // reachable states; the first in slot order is taken
        int f = 0;
#pragma unroll 1
        for (int i = 0; i < n_units; i++)
            if (M.units[i].x == tx && M.units[i].y == ty) { f = i; break; }
        o[6] = lux_ratio32_cap1(unit_space(M.units[f]), 1.0 / 100.0);
        return;
    }
    o[6] = (float)(v < 1.0 ? v : 1.0);
}

// Warp-cooperative scan of one node list for one entity: the lexicographic (squared distance, list index)
// minimum, second minimum, maximum and second maximum -- that gives both the closest and the furthest node with
// numpy's first-index tie-break, with and without the np.delete(closest) self-exclusion (:328-336).  Every lane
// scans a stride of the list keeping its two smallest / largest packed keys (d << 10 | index for the minimum,
// d << 10 | 1023 - index, + 1, for the maximum: among equal distances the FIRST index wins both ways), four
// warp reductions merge them.  Returns list indices.
struct LuxPick { int mn1, mn2, mx1, mx2; };
LUX_DEV LuxPick lux_obs_scan(const uint32_t* nodes, int n, int px, int py) {
    const int lane = lux_lane();
    unsigned lmin1 = 0xFFFFFFFFu, lmin2 = 0xFFFFFFFFu, lmax1 = 0, lmax2 = 0;
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        int xy = (int)(nodes[i] & 0xffffu);
        int dx = (xy & 0xff) - px, dy = (xy >> 8) - py;
        unsigned d = (unsigned)(dx * dx + dy * dy);
        unsigned kmin = (d << 10) | (unsigned)i, kmax = ((d << 10) | (unsigned)(1023 - i)) + 1u;
        if (kmin < lmin1) { lmin2 = lmin1; lmin1 = kmin; } else if (kmin < lmin2) lmin2 = kmin;
        if (kmax > lmax1) { lmax2 = lmax1; lmax1 = kmax; } else if (kmax > lmax2) lmax2 = kmax;
    }
    const unsigned gmin1 = lux_reduce_min_u(lmin1);
    const unsigned gmin2 = lux_reduce_min_u(lmin1 == gmin1 ? lmin2 : lmin1);
    const unsigned gmax1 = lux_reduce_max_u(lmax1);
    const unsigned gmax2 = lux_reduce_max_u(lmax1 == gmax1 ? lmax2 : lmax1);
    LuxPick P;
    P.mn1 = (int)(gmin1 & 1023u); P.mn2 = (int)(gmin2 & 1023u);
    P.mx1 = 1023 - (int)((gmax1 - 1u) & 1023u); P.mx2 = 1023 - (int)((gmax2 - 1u) & 1023u);
    return P;
}

// zero-fill `count` floats starting at p: scalar head/tail, 16-byte body
LUX_DEV void lux_zero_floats(float* p, int count) {
    const int lane = lux_lane();
    int head = (int)((16 - ((uintptr_t)p & 15)) & 15) / 4;
    if (head > count) head = count;
    if (lane < head) p[lane] = 0.0f;
    float* q = p + head;
    int body = (count - head) / 4;
#ifdef LUX_EMULATE
    for (int i = lane; i < body * 4; i += 32) q[i] = 0.0f;
#else
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* q4 = (float4*)q;
#pragma unroll 1
    for (int i = lane; i < body; i += 32) q4[i] = z;
#endif
    int done = head + body * 4;
    if (lane < count - done) p[done + lane] = 0.0f;
}

// Shared memory holds the match's lists (hdr, units, tiles, res); the grid is read in place.  Writes obs rows, entity codes and
// returns the number of entities (rows beyond `room` are dropped).
// kS: the map side when known at compile time (cell % S and cell / S are then shifts / multiplies), 0 otherwise.
template <int kS = 0>
LUX_DEV int lux_observe(LuxObsMatch& M, int team, float* g_obs, int16_t* g_ent, int room) {
    const int lane = lux_lane();
    const unsigned lt = lux_lanemask_lt();
    const LuxHeader* h = M.hdr;
    const int S = kS > 0 ? kS : M.S;
    const int n_units = h->n_units, n_tiles = h->n_tiles, n_res = h->n_res;

    // ---- node lists (agent_policy.py:197-247), order-preserving ballot compaction; resource nodes come
    // from the list entries (type + amount mirror), city tiles cost one grid load each; what later stages need
    // rides in the node entry
    int n_type[3] = {0, 0, 0};
#pragma unroll 1
    for (int base = 0; base < n_res; base += 32) {
        int r = base + lane;
        // the list entry carries the cell's type and a mirror of its amount: no grid access
        const LuxRes e = r < n_res ? M.res[r] : 0u;
        const int cell = LUX_RES_CELL(e), amount = LUX_RES_AMOUNT(e);
        const int t = amount > 0 ? LUX_RES_TYPE(e) : 0;
        const unsigned node = (unsigned)(cell % S) | ((unsigned)(cell / S) << 8) | ((unsigned)amount << 16);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            unsigned m = lux_ballot(t == k + 1);
            if (t == k + 1) M.nodes_res[k * M.r_cap + n_type[k] + lux_popc(m & lt)] = node;
            n_type[k] += lux_popc(m);
        }
    }
    int own_tiles = 0, opp_tiles = 0;
#pragma unroll 1
    for (int base = 0; base < n_tiles; base += 32) {
        int p = base + lane;
        int cell = p < n_tiles ? M.tiles[p] : 0;
        LuxCell C;
        C.flags = 0; C.tile_cd = 0; C.city_slot = 0;
        if (p < n_tiles) C = lux_ldcell(M.cells, cell);
        int own = p < n_tiles && cell_tile_team(C) == team;
        if (p < n_tiles) M.tinfo[p] = (uint8_t)((cell_tile_team(C) & 1) | ((C.tile_cd < 1) << 1));
        unsigned m = lux_ballot(own);
        if (own) M.nodes_tile[own_tiles + lux_popc(m & lt)] = (unsigned)(cell % S) | ((unsigned)(cell / S) << 8) | ((unsigned)C.city_slot << 16);
        own_tiles += lux_popc(m);
        opp_tiles += lux_popc(lux_ballot(p < n_tiles && !own));
    }
    int own_workers = 0, own_carts = 0;
#pragma unroll 1
    for (int base = 0; base < n_units; base += 32) {
        int u = base + lane;
        int mine = u < n_units && unit_team(M.units[u]) == team;
        int worker = mine && !unit_is_cart(M.units[u]);
        unsigned m = lux_ballot(worker);
        if (worker) M.nodes_work[own_workers + lux_popc(m & lt)] = (unsigned)M.units[u].x | ((unsigned)M.units[u].y << 8);
        own_workers += lux_popc(m);
        own_carts += lux_popc(lux_ballot(mine && !worker));
    }
    lux_sync();

    const int night = (h->turn % LUX_CYCLE_LENGTH) >= LUX_DAY_LENGTH;
    const float f_turn = lux_ratio32(h->turn, 1.0 / 360.0);
    const float f_own_tiles = lux_ratio32(own_tiles, 1.0 / 30.0), f_opp_tiles = lux_ratio32(opp_tiles, 1.0 / 30.0);
    const float f_workers = lux_ratio32(own_workers, 1.0 / 30.0), f_carts = lux_ratio32(own_carts, 1.0 / 30.0);
    const int rp = h->research[team];
    const float f_rp = lux_ratio32(rp, 1.0 / 200.0);

    // ---- entities: the team's units with can_act, then its tiles with can_act, 32 per pass
    int n_ent = 0;
    const int total = n_units + n_tiles;
#pragma unroll 1
    for (int base = 0; base < total; base += 32) {
        int e = base + lane;
        int is_unit = e < n_units;
        int px = 0, py = 0, kind = 0, code = 0, valid = 0;
        if (is_unit) {
            const LuxUnit& U = M.units[e];
            valid = unit_team(U) == team && U.cd_q < 4;
            px = U.x; py = U.y; kind = unit_is_cart(U) ? 1 : 0; code = e;
        } else if (e < total) {
            int p = e - n_units;
            int cell = M.tiles[p];
            const int ti = M.tinfo[p];
            valid = (ti & 1) == team && (ti & 2) != 0;
            px = cell % S; py = cell / S; kind = 2; code = 0x4000 | p;
        }
        unsigned vm = lux_ballot(valid);
        int nv = lux_popc(vm);
        if (nv == 0) continue;
        int row = n_ent + lux_popc(vm & lt);
        int rows_out = n_ent + nv <= room ? nv : (room > n_ent ? room - n_ent : 0);
        // the rows of this pass are zero-filled by the whole warp (coalesced, 16-byte stores) ...
        lux_zero_floats(g_obs + (size_t)n_ent * LUX_OBS_DIM, rows_out * LUX_OBS_DIM);
        lux_sync();
        // ... then the warp takes the entities of this pass one at a time: a typical match has one or two
        // actionable entities, so lanes-over-entities would leave the list scans to one or two lanes
        unsigned todo = vm;
#pragma unroll 1
        while (todo) {
            const int src = lux_ffs(todo) - 1;
            todo &= todo - 1;
            const int e_row = lux_shfl(row, src);
            if (e_row >= room) continue;                              // warp-uniform
            const int e_px = lux_shfl(px, src), e_py = lux_shfl(py, src), e_kind = lux_shfl(kind, src);
            const int e_code = lux_shfl(code, src), e_idx = base + src;
            float* o = g_obs + (size_t)e_row * LUX_OBS_DIM;
            int alone = 1;
            if (e_kind == 0) {                       // len(cell.units) <= 1 (agent_policy.py:330)
                int cnt = 0;
#pragma unroll 1
                for (int cb = 0; cb < n_units; cb += 32) {
                    int v = cb + lane;
                    cnt += lux_popc(lux_ballot(v < n_units && M.units[v].x == e_px && M.units[v].y == e_py));
                }
                alone = cnt <= 1;
            }
#pragma unroll 1
            for (int key = 0; key < 5; key++) {
                const uint32_t* nodes = key < 3 ? M.nodes_res + key * M.r_cap : key == 3 ? M.nodes_tile : M.nodes_work;
                const int n = key < 3 ? n_type[key] : key == 3 ? own_tiles : own_workers;
                if (n == 0) continue;                                 // key not in object_nodes
                const int exclude = (key == 3 && e_kind == 2) || (key == 4 && e_kind == 0 && alone);
                const LuxPick P = lux_obs_scan(nodes, n, e_px, e_py);
                // lane `key` writes the closest block, lane 5 + key the furthest one
                if (lane == key || lane == 5 + key) {
                    const int furthest = lane >= 5;
                    float* blk = o + 3 + 35 * furthest + 7 * key;
                    if (exclude && n == 1) {
                        blk[5] = 1.0f;                                // nothing left after np.delete
                    } else {
                        int pick;
                        if (exclude) pick = !furthest ? P.mn2 : (P.mx1 == P.mn1 ? P.mx2 : P.mx1);
                        else pick = furthest ? P.mx1 : P.mn1;
                        lux_obs_emit(M, blk, nodes[pick], e_px, e_py, key, n_units);
                    }
                }
            }
            if (lane == 10) {
                o[e_kind] = 1.0f;
                if (e_kind != 2) o[73] = lux_ratio32(unit_space(M.units[e_idx]), 1.0 / 100.0);
                o[74] = night ? 1.0f : 0.0f;
                o[75] = f_turn;
                o[76] = f_own_tiles; o[77] = f_opp_tiles;
                o[78] = f_workers; o[79] = f_workers;       // agent_policy.py:214-215: the "opponent" lists hold the own units
                o[80] = f_carts; o[81] = f_carts;
                o[82] = f_rp;
                o[83] = rp >= LUX_RESEARCH_COAL ? 1.0f : 0.0f;
                o[84] = rp >= LUX_RESEARCH_URANIUM ? 1.0f : 0.0f;
                g_ent[e_row] = (int16_t)e_code;
            }
        }
        n_ent += nv;
        lux_sync();
    }
    return n_ent;
}
