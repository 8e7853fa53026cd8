/*
 * lux_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A sequential CPU restatement of the reference turn engine, operating on the packed
 * per-match state of include/lux_b200.h.  It is the checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg) and
 * is never linked into or called from the product.  It deliberately keeps the
 * reference's own data structures and iteration orders (per-team unit lists, ordered
 * city-cell lists, per-cell request lists, the recursive collision revert) rather than
 * the parallel formulation used on the GPU, so that the two are independent derivations.
 *
 * Pinned (tests/test_oracle_golden.py) against golden vectors produced by the unmodified
 * reference (oracle/make_golden.py): the 7 Kaggle replays of
 * luxai2021/tests/replays_for_test and seeded random-action matches on 12/16/24/32 maps.
 *
 * Reference citations are to /root/reference/luxai2021/game/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lux_b200.h"

#define MAXCELLS 1024
#define MAXU 1024
#define MAXC 1024

/* game_constants.json:18-57 */
enum {
    DAY_LENGTH = 30, NIGHT_LENGTH = 10, MAX_DAYS = 360,
    UPKEEP_CITY = 23, UPKEEP_WORKER = 4, UPKEEP_CART = 10,
    MAX_WOOD = 500, CITY_BUILD_COST = 100, ADJ_BONUS = 5,
    CAP_WORKER = 100, CAP_CART = 2000,
    RESEARCH_COAL = 50, RESEARCH_URANIUM = 200,
    CITY_ACTION_COOLDOWN = 10, CD_CART = 3, CD_WORKER = 2,
    MAX_ROAD_Q = 24, ROAD_DEV_Q = 3, PILLAGE_Q = 2
};
static const int RATE[4] = {0, 20, 5, 2};      /* WORKER_COLLECTION_RATE by resource code */
static const int FUEL[4] = {0, 1, 10, 40};     /* RESOURCE_TO_FUEL_RATE */
static const double WOOD_GROWTH_RATE = 1.025;

typedef struct {
    int id, x, y, team, cart, cd_q, cargo[4]; /* cargo[1..3] wood/coal/uranium */
    uint32_t word;                             /* raw action word of this turn */
    int act;                                   /* validated/held action kind, 0 none */
    int dir;                                   /* for a held move */
    int alive;
} OUnit;

typedef struct {
    int id, team, fuel;
    int ncells;
    int cells[MAXCELLS];                       /* city_cells, append order */
    int alive;
} OCity;

typedef struct {
    int S;
    LuxHeader* h;
    LuxCell* cells;
    /* per-team unit lists: teamStates[t]["units"] dict order */
    OUnit* tu[2];
    int ntu[2];
    OCity* cities;                             /* Game.cities dict order (dead ones flagged) */
    int ncities;
    int cell_city[MAXCELLS];                   /* index into cities[] for tile cells */
    int u_cap, c_cap, t_cap;
} Game;

static int tile_team(const LuxCell* c) { return ((c->flags >> 2) & 3) - 1; }
static int is_tile(const LuxCell* c) { return ((c->flags >> 2) & 3) != 0; }
static int res_type(const LuxCell* c) { return c->flags & 3; }
static int has_resource(const LuxCell* c) { return c->amount > 0; } /* cell.py:44-49 */
static int cell_adj(const LuxCell* c) { return (c->flags >> 4) & 7; }
static void set_adj(LuxCell* c, int a) { c->flags = (uint8_t)((c->flags & 0x8F) | (a << 4)); }

static const int DX[5] = {0, 0, 1, 0, -1};    /* c n e s w : position.py:36-46 */
static const int DY[5] = {0, -1, 0, 1, 0};

static int space_left(const OUnit* u) {       /* unit.py:43-52 */
    return (u->cart ? CAP_CART : CAP_WORKER) - (u->cargo[1] + u->cargo[2] + u->cargo[3]);
}

/* game_map.py:484-508 : N, E, S, W */
static int adjacent_cells(int S, int idx, int out[4]) {
    int x = idx % S, y = idx / S, n = 0;
    if (y > 0) out[n++] = idx - S;
    if (x < S - 1) out[n++] = idx + 1;
    if (y < S - 1) out[n++] = idx + S;
    if (x > 0) out[n++] = idx - 1;
    return n;
}

/* ---------------------------------------------------------------- unpack / pack */

static void unpack(Game* g, LuxHeader* h, LuxCell* cells, LuxUnit* units, LuxCity* cities,
                   uint16_t* tiles, const uint32_t* unit_actions) {
    g->S = h->size;
    g->h = h;
    g->cells = cells;
    g->ntu[0] = g->ntu[1] = 0;
    for (int i = 0; i < h->n_units; i++) {
        const LuxUnit* p = &units[i];
        int team = p->team_type & 1;
        OUnit* u = &g->tu[team][g->ntu[team]++];
        memset(u, 0, sizeof *u);
        u->id = p->id; u->x = p->x; u->y = p->y; u->team = team; u->cart = (p->team_type >> 1) & 1;
        u->cd_q = p->cd_q; u->cargo[1] = p->wood; u->cargo[2] = p->coal; u->cargo[3] = p->uranium;
        u->word = unit_actions ? unit_actions[i] : 0;
        u->alive = 1;
    }
    g->ncities = h->n_cities;
    for (int s = 0; s < h->n_cities; s++) {
        OCity* c = &g->cities[s];
        c->id = cities[s].id; c->team = cities[s].team; c->fuel = cities[s].fuel;
        c->ncells = 0; c->alive = 1;
    }
    /* tiles[] is cities x city_cells order, keys are dense, so appending in list order
       rebuilds every city_cells list */
    for (int p = 0; p < h->n_tiles; p++) {
        int idx = tiles[p];
        OCity* c = &g->cities[cells[idx].city_slot];
        c->cells[c->ncells++] = idx;
        g->cell_city[idx] = cells[idx].city_slot;
    }
}

static int team_tile_count_packed(Game* g, int team) {
    int n = 0;
    for (int s = 0; s < g->ncities; s++)
        if (g->cities[s].alive && g->cities[s].team == team) n += g->cities[s].ncells;
    return n;
}

static void pack(Game* g, LuxUnit* units, LuxCity* cities, uint16_t* tiles) {
    LuxHeader* h = g->h;
    int n = 0;
    for (int t = 0; t < 2; t++) {
        int cnt = 0;
        for (int i = 0; i < g->ntu[t]; i++) {
            OUnit* u = &g->tu[t][i];
            if (!u->alive) continue;
            LuxUnit* p = &units[n++];
            memset(p, 0, sizeof *p);
            p->id = u->id; p->x = (uint8_t)u->x; p->y = (uint8_t)u->y;
            p->team_type = (uint8_t)(u->team | (u->cart << 1)); p->cd_q = (uint8_t)u->cd_q;
            p->wood = (uint16_t)u->cargo[1]; p->coal = (uint16_t)u->cargo[2]; p->uranium = (uint16_t)u->cargo[3];
            cnt++;
        }
        h->n_team_units[t] = (uint16_t)cnt;
    }
    h->n_units = (uint16_t)n;
    int slot = 0, nt = 0;
    for (int s = 0; s < g->ncities; s++) {
        OCity* c = &g->cities[s];
        if (!c->alive) continue;
        LuxCity* p = &cities[slot];
        memset(p, 0, sizeof *p);
        p->id = c->id; p->fuel = c->fuel; p->team = (uint8_t)c->team; p->ntiles = (uint16_t)c->ncells;
        int adjsum = 0;
        for (int k = 0; k < c->ncells; k++) {
            LuxCell* cc = &g->cells[c->cells[k]];
            cc->city_slot = (uint8_t)slot;
            cc->key = (uint16_t)k;
            adjsum += cell_adj(cc);
            if (nt < g->t_cap) tiles[nt] = (uint16_t)c->cells[k];
            nt++;
        }
        p->adjsum = (uint16_t)adjsum;
        slot++;
    }
    h->n_cities = (uint16_t)slot;
    h->n_tiles = (uint16_t)nt;
    h->n_team_tiles[0] = (uint16_t)team_tile_count_packed(g, 0);
    h->n_team_tiles[1] = (uint16_t)team_tile_count_packed(g, 1);
}

/* ---------------------------------------------------------------- helpers */

static OUnit* find_unit(Game* g, int team, int id) {   /* game.py:1032-1037 */
    for (int i = 0; i < g->ntu[team]; i++)
        if (g->tu[team][i].alive && g->tu[team][i].id == id) return &g->tu[team][i];
    return NULL;
}

static int team_tile_count(Game* g, int team) {         /* game.py:725-735 */
    int n = 0;
    for (int s = 0; s < g->ncities; s++)
        if (g->cities[s].alive && g->cities[s].team == team) n += g->cities[s].ncells;
    return n;
}

static int live_units(Game* g, int team) {
    int n = 0;
    for (int i = 0; i < g->ntu[team]; i++) n += g->tu[team][i].alive;
    return n;
}

static int is_night(const LuxHeader* h) {                /* game.py:1197-1204 */
    return (h->turn % (DAY_LENGTH + NIGHT_LENGTH)) >= DAY_LENGTH;
}

/* game.py:744-796 spawn_worker / spawn_cart (engine path: no explicit id) */
static void spawn_unit(Game* g, int team, int x, int y, int cart) {
    LuxHeader* h = g->h;
    if (live_units(g, 0) + live_units(g, 1) >= g->u_cap || g->ntu[team] >= MAXU) {
        h->status |= LUX_ST_UNIT_OVERFLOW;
        return;
    }
    OUnit* u = &g->tu[team][g->ntu[team]++];
    memset(u, 0, sizeof *u);
    u->id = ++h->unit_id_count;
    u->x = x; u->y = y; u->team = team; u->cart = cart; u->alive = 1;
    if (cart) h->carts_built[team]++; else h->workers_built[team]++;
}

/* game.py:798-854 spawn_city_tile */
static void spawn_city_tile(Game* g, int team, int idx) {
    LuxHeader* h = g->h;
    LuxCell* cell = &g->cells[idx];
    int adj[4], nadj = adjacent_cells(g->S, idx, adj);
    int same[4], nsame = 0, found[4], nfound = 0;
    for (int k = 0; k < nadj; k++) {
        LuxCell* c2 = &g->cells[adj[k]];
        if (is_tile(c2) && tile_team(c2) == team) {
            same[nsame++] = adj[k];
            int cid = g->cell_city[adj[k]], seen = 0;
            for (int f = 0; f < nfound; f++) seen |= (found[f] == cid);
            if (!seen) found[nfound++] = cid;
        }
    }
    int live_tiles = team_tile_count(g, 0) + team_tile_count(g, 1);
    if (live_tiles >= g->t_cap) { h->status |= LUX_ST_TILE_OVERFLOW; return; }
    if (nsame == 0) {
        int live_cities = 0;
        for (int s = 0; s < g->ncities; s++) live_cities += g->cities[s].alive;
        if (live_cities >= g->c_cap || g->ncities >= MAXC) { h->status |= LUX_ST_CITY_OVERFLOW; return; }
        OCity* c = &g->cities[g->ncities];
        c->id = ++h->city_id_count; c->team = team; c->fuel = 0; c->ncells = 0; c->alive = 1;
        cell->flags = (uint8_t)((cell->flags & 0x83) | ((team + 1) << 2));   /* adjacent_city_tiles = 0 */
        cell->tile_cd = 0;
        c->cells[c->ncells++] = idx;
        g->cell_city[idx] = g->ncities;
        g->ncities++;
        return;
    }
    int cid = g->cell_city[same[0]];
    OCity* city = &g->cities[cid];
    cell->flags = (uint8_t)((cell->flags & 0x83) | ((team + 1) << 2));
    cell->tile_cd = 0;
    set_adj(cell, nsame);
    for (int k = 0; k < nsame; k++) set_adj(&g->cells[same[k]], cell_adj(&g->cells[same[k]]) + 1);
    city->cells[city->ncells++] = idx;
    g->cell_city[idx] = cid;
    for (int f = 0; f < nfound; f++) {
        if (found[f] == cid) continue;
        OCity* old = &g->cities[found[f]];
        for (int k = 0; k < old->ncells; k++) {
            g->cell_city[old->cells[k]] = cid;
            city->cells[city->ncells++] = old->cells[k];
        }
        city->fuel += old->fuel;
        old->alive = 0;
        old->ncells = 0;
    }
}

/* ---------------------------------------------------------------- movement (game.py:1104-1195) */

typedef struct {
    Game* g;
    int nmoves;
    OUnit* mover[MAXU];
    int target[MAXU];
    int next[MAXU];               /* per-cell move lists, insertion order */
    int present[MAXCELLS];        /* key exists in cells_to_actions_to_there */
    int head[MAXCELLS], tail[MAXCELLS], nlist[MAXCELLS];
} MoveCtx;

static void revert_action(MoveCtx* m, int a) {
    OUnit* u = m->mover[a];
    int orig = u->y * m->g->S + u->x;
    if (!is_tile(&m->g->cells[orig])) {
        if (m->present[orig]) {
            m->present[orig] = 0;                                  /* pop(original_cell) */
            for (int k = m->head[orig]; k >= 0; k = m->next[k]) revert_action(m, k);
        }
    }
}

static void handle_movement(Game* g) {
    static __thread MoveCtx ctx;
    MoveCtx* m = &ctx;
    int S = g->S, ncell = S * S;
    m->g = g;
    m->nmoves = 0;
    memset(m->present, 0, sizeof(int) * ncell);
    memset(m->nlist, 0, sizeof(int) * ncell);
    int order[MAXCELLS], norder = 0;     /* key insertion order */
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < g->ntu[t]; i++) {
            OUnit* u = &g->tu[t][i];
            if (u->act != LUX_UA_MOVE) continue;
            int a = m->nmoves++;
            m->mover[a] = u;
            int tgt = (u->y + DY[u->dir]) * S + (u->x + DX[u->dir]);
            m->target[a] = tgt;
            m->next[a] = -1;
            if (!m->present[tgt]) { m->present[tgt] = 1; order[norder++] = tgt; m->head[tgt] = a; }
            else m->next[m->tail[tgt]] = a;
            m->tail[tgt] = a;
            m->nlist[tgt]++;
        }
    for (int o = 0; o < norder; o++) {
        int cell = order[o];
        if (!m->present[cell]) continue;
        static __thread int revert[MAXU];
        int nrev = 0;
        const LuxCell* c = &g->cells[cell];
        if (m->nlist[cell] > 1) {
            if (!is_tile(c))
                for (int k = m->head[cell]; k >= 0; k = m->next[k]) revert[nrev++] = k;
        } else if (m->nlist[cell] == 1) {
            if (!is_tile(c)) {
                int count = 0, still = 1;
                for (int t = 0; t < 2; t++)
                    for (int i = 0; i < g->ntu[t]; i++) {
                        OUnit* u = &g->tu[t][i];
                        if (u->alive && u->y * S + u->x == cell) {
                            count++;
                            if (u->act == LUX_UA_MOVE) still = 0;  /* id in moving_units */
                        }
                    }
                if (count == 1 && still) revert[nrev++] = m->head[cell];
            }
        }
        for (int k = 0; k < nrev; k++) revert_action(m, revert[k]);
        for (int k = 0; k < nrev; k++) m->present[m->target[revert[k]]] = 0;
    }
    /* pruned actions; a CENTER survivor is not given to the unit (game.py:462-465) */
    for (int a = 0; a < m->nmoves; a++) {
        OUnit* u = m->mover[a];
        if (!m->present[m->target[a]] || u->dir == LUX_DIR_CENTER) u->act = LUX_UA_NONE;
    }
}

/* ---------------------------------------------------------------- collection (game.py:868-1001) */

typedef struct { int from, amount, team; OUnit* worker; int city; } Req;

static void handle_resource_type(Game* g, int rtype) {
    static __thread Req reqs[MAXCELLS][64];
    static __thread int nreq[MAXCELLS];
    int S = g->S, ncell = S * S;
    int order[MAXCELLS], norder = 0;
    memset(nreq, 0, sizeof(int) * ncell);
    LuxHeader* h = g->h;
    for (int team = 0; team < 2; team++) {
        int researched = rtype == LUX_RES_WOOD ? 1
                       : rtype == LUX_RES_COAL ? h->research[team] >= RESEARCH_COAL
                                               : h->research[team] >= RESEARCH_URANIUM;
        if (!researched) continue;
        for (int i = 0; i < g->ntu[team]; i++) {
            OUnit* u = &g->tu[team][i];
            if (!u->alive || u->cart) continue;
            int ucell = u->y * S + u->x;
            int cand[5], ncand = 0, adj[4];
            cand[ncand++] = ucell;
            int na = adjacent_cells(S, ucell, adj);
            for (int k = 0; k < na; k++) cand[ncand++] = adj[k];
            int minable[5], nmin = 0;
            for (int k = 0; k < ncand; k++) {
                const LuxCell* c = &g->cells[cand[k]];
                if (has_resource(c) && res_type(c) == rtype) minable[nmin++] = cand[k];
            }
            if (nmin == 0) continue;
            int space = space_left(u);
            int mine = (space + nmin - 1) / nmin;          /* math.ceil(space / len(minable)) */
            if (mine > RATE[rtype]) mine = RATE[rtype];
            int on_tile = is_tile(&g->cells[ucell]);
            for (int k = 0; k < nmin; k++) {
                int c = minable[k];
                Req r = {ucell, mine, team, on_tile ? NULL : u, on_tile ? g->cell_city[ucell] : -1};
                int dup = 0;                                 /* ResourceRequest.__eq__ :920-931 */
                for (int q = 0; q < nreq[c]; q++) {
                    Req* o = &reqs[c][q];
                    int wid = r.worker ? r.worker->id : -1, owid = o->worker ? o->worker->id : -1;
                    if (o->from == r.from && wid == owid && o->amount == r.amount && o->city == r.city) dup = 1;
                }
                if (!dup && nreq[c] < 64) {
                    if (nreq[c] == 0) order[norder++] = c;
                    reqs[c][nreq[c]++] = r;
                }
            }
        }
    }
    for (int o = 0; o < norder; o++) {                       /* resolve_resource_requests :969-1001 */
        int c = order[o];
        int left = g->cells[c].amount;
        Req* live[64];
        int n = 0;
        for (int q = 0; q < nreq[c]; q++) live[n++] = &reqs[c][q];
        for (;;) {
            int sum = 0, mn = 1 << 30;
            for (int q = 0; q < n; q++) { sum += live[q]->amount; if (live[q]->amount < mn) mn = live[q]->amount; }
            if (!(n > 0 && sum > 0 && left > 0)) break;
            int fill = left / n;
            if (mn < fill) fill = mn;
            for (int q = 0; q < n; q++) {
                Req* r = live[q];
                if (r->city >= 0) {
                    OCity* city = &g->cities[r->city];
                    h->collected[city->team][rtype - 1] += fill;
                    city->fuel += fill * FUEL[rtype];
                } else {
                    int give = space_left(r->worker);
                    if (fill < give) give = fill;
                    h->collected[r->worker->team][rtype - 1] += give;
                    r->worker->cargo[rtype] += give;
                }
                r->amount -= fill;
            }
            left -= fill * n;
            if (left < n) left = 0;
            int m2 = 0;
            for (int q = 0; q < n; q++) if (live[q]->amount > 0) live[m2++] = live[q];
            n = m2;
        }
        g->cells[c].amount = (uint16_t)left;
    }
}

/* ---------------------------------------------------------------- the turn (game.py:390-535) */

static void oracle_turn(Game* g, const uint8_t* tile_actions, const uint16_t* tiles_in, int n_tiles_in,
                        LuxRes* res, int n_res) {
    LuxHeader* h = g->h;
    int S = g->S;
    int night = is_night(h);
    int mult = night ? 2 : 1;

    /* -- validate_command over [tile actions..., unit actions...] (game.py:406-425, :646-723) */
    int acc[2] = {0, 0};
    int team_units[2] = {live_units(g, 0), live_units(g, 1)};
    int team_tiles[2] = {team_tile_count(g, 0), team_tile_count(g, 1)};
    static __thread uint8_t tile_act[MAXCELLS];
    for (int p = 0; p < n_tiles_in; p++) {
        int code = tile_actions ? tile_actions[p] : 0;
        const LuxCell* c = &g->cells[tiles_in[p]];
        int team = tile_team(c);
        tile_act[p] = 0;
        if (code == LUX_TA_BUILD_WORKER || code == LUX_TA_BUILD_CART) {
            if (!(c->tile_cd < 1)) continue;                             /* :703 */
            if (team_units[team] + acc[team] >= team_tiles[team]) continue; /* :705-710, :725-742 */
            acc[team]++;
            tile_act[p] = (uint8_t)code;
        } else if (code == LUX_TA_RESEARCH) {
            tile_act[p] = (uint8_t)code;                                  /* :721-723 no check */
        }
    }
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < g->ntu[t]; i++) {
            OUnit* u = &g->tu[t][i];
            int kind = u->word & 7;
            u->act = LUX_UA_NONE;
            const LuxCell* c = &g->cells[u->y * S + u->x];
            if (kind == LUX_UA_BCITY) {                                   /* :656-675 */
                if (is_tile(c) || has_resource(c)) continue;
                if (!(u->cd_q < 4)) continue;
                if (u->cargo[1] + u->cargo[2] + u->cargo[3] < CITY_BUILD_COST) continue;
                u->act = kind;
            } else if (kind == LUX_UA_MOVE) {                             /* :676-695 */
                int dir = (u->word >> 3) & 7;
                if (!(u->cd_q < 4)) continue;
                if (dir > 4) continue;
                if (dir != LUX_DIR_CENTER) {
                    int nx = u->x + DX[dir], ny = u->y + DY[dir];
                    if (nx < 0 || ny < 0 || nx >= S || ny >= S) continue;
                    const LuxCell* tc = &g->cells[ny * S + nx];
                    if (is_tile(tc) && tile_team(tc) != u->team) continue;
                }
                u->act = kind;
                u->dir = dir;
            } else if (kind == LUX_UA_PILLAGE || kind == LUX_UA_TRANSFER) {
                u->act = kind;                                            /* :721-723 */
            }
        }

    /* -- collisions (:455-465) */
    handle_movement(g);

    /* -- city tile turns, cities x city_cells order (:468-475, city.py:96-120) */
    for (int p = 0; p < n_tiles_in; p++) {
        int idx = tiles_in[p];
        LuxCell* c = &g->cells[idx];
        int act = tile_act[p];
        if (act == LUX_TA_BUILD_CART) { spawn_unit(g, tile_team(c), idx % S, idx / S, 1); c->tile_cd = CITY_ACTION_COOLDOWN; }
        else if (act == LUX_TA_BUILD_WORKER) { spawn_unit(g, tile_team(c), idx % S, idx / S, 0); c->tile_cd = CITY_ACTION_COOLDOWN; }
        else if (act == LUX_TA_RESEARCH) { c->tile_cd = CITY_ACTION_COOLDOWN; h->research[tile_team(c)]++; }
        if (c->tile_cd > 0) c->tile_cd--;
    }

    /* -- unit turns, team A then B (:477-485, unit.py:162-197, :234-269) */
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < g->ntu[t]; i++) {
            OUnit* u = &g->tu[t][i];
            if (!u->alive) continue;
            int failed = 0;
            if (u->act != LUX_UA_NONE) {
                if (u->act == LUX_UA_MOVE) {                               /* move_unit :856-866 */
                    u->x += DX[u->dir]; u->y += DY[u->dir];
                } else if (u->act == LUX_UA_TRANSFER) {                    /* transfer_resources :1039-1057 */
                    int r = ((u->word >> 3) & 3) + 1, amount = (u->word >> 6) & 2047, dest = (int)(u->word >> 17);
                    OUnit* d = find_unit(g, t, dest);
                    if (d == NULL || r > 3) failed = 1;                    /* KeyError: swallowed at :480-485 */
                    else {
                        int amt = amount;
                        if (u->cargo[r] < amt) amt = u->cargo[r];
                        if (space_left(d) < amt) amt = space_left(d);
                        u->cargo[r] -= amt;
                        d->cargo[r] += amt;
                    }
                } else if (u->act == LUX_UA_BCITY && !u->cart) {
                    spawn_city_tile(g, t, u->y * S + u->x);
                    int spent = 0;                                         /* expend_resources_for_city unit.py:146-160 */
                    for (int r = 1; r <= 3; r++) {
                        if (spent + u->cargo[r] > CITY_BUILD_COST) { u->cargo[r] -= CITY_BUILD_COST - spent; break; }
                        spent += u->cargo[r];
                        u->cargo[r] = 0;
                    }
                } else if (u->act == LUX_UA_PILLAGE && !u->cart) {
                    LuxCell* c = &g->cells[u->y * S + u->x];
                    c->road_q = (uint8_t)(c->road_q > PILLAGE_Q ? c->road_q - PILLAGE_Q : 0);
                }
                if (!failed) u->cd_q += 4 * (u->cart ? CD_CART : CD_WORKER) * mult;
            }
            u->act = LUX_UA_NONE;
            if (u->cart && !failed) {                                       /* unit.py:259-269 */
                LuxCell* c = &g->cells[u->y * S + u->x];
                if (!is_tile(c) && c->road_q < MAX_ROAD_Q) {
                    int r = c->road_q + ROAD_DEV_Q;
                    c->road_q = (uint8_t)(r > MAX_ROAD_Q ? MAX_ROAD_Q : r);
                    h->roads_built_q[t] += ROAD_DEV_Q;
                }
            }
        }

    /* -- distribute_all_resources (:868-880): uranium, coal, wood */
    handle_resource_type(g, LUX_RES_URANIUM);
    handle_resource_type(g, LUX_RES_COAL);
    handle_resource_type(g, LUX_RES_WOOD);

    /* -- handle_resource_deposit (:1003-1023) */
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < g->ntu[t]; i++) {
            OUnit* u = &g->tu[t][i];
            if (!u->alive) continue;
            int idx = u->y * S + u->x;
            const LuxCell* c = &g->cells[idx];
            if (is_tile(c) && tile_team(c) == u->team) {
                int fuel = u->cargo[1] * FUEL[1] + u->cargo[2] * FUEL[2] + u->cargo[3] * FUEL[3];
                g->cities[g->cell_city[idx]].fuel += fuel;
                h->fuel_generated[u->team] += fuel;
                u->cargo[1] = u->cargo[2] = u->cargo[3] = 0;
            }
        }

    /* -- handle_night (:537-558) */
    if (night) {
        for (int s = 0; s < g->ncities; s++) {
            OCity* city = &g->cities[s];
            if (!city->alive) continue;
            int bonus = 0;
            for (int k = 0; k < city->ncells; k++) bonus += cell_adj(&g->cells[city->cells[k]]) * ADJ_BONUS;
            int upkeep = city->ncells * UPKEEP_CITY - bonus;               /* city.py:34-50 */
            if (city->fuel < upkeep) {                                      /* destroy_city :1059-1070 */
                for (int k = 0; k < city->ncells; k++) {
                    LuxCell* c = &g->cells[city->cells[k]];
                    c->flags &= 0x03;
                    c->tile_cd = 0; c->city_slot = 0; c->key = 0;
                    c->road_q = 0;
                }
                city->alive = 0;
                city->ncells = 0;
            } else {
                city->fuel -= upkeep;
            }
        }
        for (int t = 0; t < 2; t++)
            for (int i = 0; i < g->ntu[t]; i++) {
                OUnit* u = &g->tu[t][i];
                if (!u->alive) continue;
                if (is_tile(&g->cells[u->y * S + u->x])) continue;
                int need = u->cart ? UPKEEP_CART : UPKEEP_WORKER;          /* spend_fuel_to_survive unit.py:64-98 */
                for (int r = 1; r <= 3 && need > 0; r++) {
                    int want = (need + FUEL[r] - 1) / FUEL[r];
                    int used = u->cargo[r] < want ? u->cargo[r] : want;
                    need -= used * FUEL[r];
                    u->cargo[r] -= used;
                }
                if (need > 0) u->alive = 0;                                  /* destroy_unit :1072-1079 */
            }
    }

    /* -- depleted cells leave the resource list (:498-510), trees regrow (:1081-1102) */
    for (int k = 0; k < n_res; k++) {
        LuxCell* c = &g->cells[LUX_RES_CELL(res[k])];
        if (c->amount == 0) c->flags &= 0xFC;
        else if (res_type(c) == LUX_RES_WOOD && c->amount < MAX_WOOD) {
            double grown = c->amount * WOOD_GROWTH_RATE;
            if (grown > (double)MAX_WOOD) grown = (double)MAX_WOOD;
            c->amount = (uint16_t)ceil(grown);
        }
        /* the list entry mirrors the cell's amount (include/lux_b200.h LuxRes); the grid is the oracle's truth */
        res[k] = LUX_RES_MAKE((uint32_t)LUX_RES_CELL(res[k]), (uint32_t)LUX_RES_TYPE(res[k]), c->amount);
    }

    /* -- match_over (:570-592), before the turn counter moves (:515-517) */
    int over = h->turn >= MAX_DAYS - 1;
    int ncity[2] = {0, 0};
    for (int s = 0; s < g->ncities; s++) if (g->cities[s].alive) ncity[g->cities[s].team]++;
    for (int t = 0; t < 2; t++) if (live_units(g, t) + ncity[t] == 0) over = 1;
    h->turn++;

    /* -- run_cooldowns (:560-568) */
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < g->ntu[t]; i++) {
            OUnit* u = &g->tu[t][i];
            if (!u->alive) continue;
            const LuxCell* c = &g->cells[u->y * S + u->x];
            int road = is_tile(c) ? MAX_ROAD_Q : c->road_q;               /* cell.py:77-85 */
            int cd = u->cd_q - road - 4;
            u->cd_q = cd > 0 ? cd : 0;
        }

    if (over) {
        h->done = 1;
        int tiles_t[2] = {team_tile_count(g, 0), team_tile_count(g, 1)};  /* get_winning_team :594-634 */
        int units_t[2] = {live_units(g, 0), live_units(g, 1)};
        if (tiles_t[0] != tiles_t[1]) h->winner = tiles_t[0] > tiles_t[1] ? LUX_WIN_A : LUX_WIN_B;
        else if (units_t[0] != units_t[1]) h->winner = units_t[0] > units_t[1] ? LUX_WIN_A : LUX_WIN_B;
        else if (h->fuel_generated[0] != h->fuel_generated[1])
            h->winner = h->fuel_generated[0] > h->fuel_generated[1] ? LUX_WIN_A : LUX_WIN_B;
        else h->winner = LUX_WIN_TIE;                                      /* :632 host coin flip */
    }
}

/* ---------------------------------------------------------------- public entry points */

typedef struct { OUnit* tu[2]; OCity* cities; } Scratch;

static Scratch* scratch_get(void) {
    static __thread Scratch s;
    if (!s.cities) {
        s.tu[0] = (OUnit*)malloc(sizeof(OUnit) * MAXU);
        s.tu[1] = (OUnit*)malloc(sizeof(OUnit) * MAXU);
        s.cities = (OCity*)malloc(sizeof(OCity) * MAXC);
    }
    return &s;
}

/* One turn of one match, in place.  Returns 0, or LUX_E_ARG. */
int lux_oracle_step(LuxHeader* hdr, LuxCell* cells, LuxUnit* units, LuxCity* cities, uint16_t* tiles,
                    LuxRes* res, const uint32_t* unit_actions, const uint8_t* tile_actions,
                    int u_cap, int c_cap, int t_cap) {
    if (!hdr || !cells || !units || !cities || !tiles || !res) return LUX_E_ARG;
    if (hdr->size < 2 || hdr->size > 32 || u_cap > MAXU || c_cap > 256 || t_cap > MAXCELLS) return LUX_E_ARG;
    if (hdr->done) return 0;
    Scratch* s = scratch_get();
    static __thread Game g;
    g.tu[0] = s->tu[0]; g.tu[1] = s->tu[1]; g.cities = s->cities;
    g.u_cap = u_cap; g.c_cap = c_cap; g.t_cap = t_cap;
    unpack(&g, hdr, cells, units, cities, tiles, unit_actions);
    uint16_t tiles_in[MAXCELLS];
    int n_tiles_in = hdr->n_tiles;
    memcpy(tiles_in, tiles, sizeof(uint16_t) * n_tiles_in);
    oracle_turn(&g, tile_actions, tiles_in, n_tiles_in, res, hdr->n_res);
    pack(&g, units, cities, tiles);
    return 0;
}

/* Batch form over host memory with the section strides of LuxBatch (one thread). */
int lux_oracle_step_batch(const LuxBatch* b, const uint32_t* unit_actions, const uint8_t* tile_actions,
                          int first, int count) {
    if (!b) return LUX_E_ARG;
    int ncell = b->size * b->size;
    for (int m = first; m < first + count && m < b->n_matches; m++) {
        int rc = lux_oracle_step(&b->hdr[m], b->cells + (size_t)m * ncell, b->units + (size_t)m * b->u_cap,
                                 b->cities + (size_t)m * b->c_cap, b->tiles + (size_t)m * b->t_cap,
                                 b->res + (size_t)m * b->r_cap,
                                 unit_actions ? unit_actions + (size_t)m * b->u_cap : NULL,
                                 tile_actions ? tile_actions + (size_t)m * b->t_cap : NULL,
                                 b->u_cap, b->c_cap, b->t_cap);
        if (rc) return rc;
    }
    return 0;
}

/* ---------------------------------------------------------------- canonical action stream */

static uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}

uint64_t lux_oracle_hash(uint64_t seed, uint64_t match, uint64_t turn, uint64_t kind, uint64_t key) {
    const uint64_t G = 0x9E3779B97F4A7C15ULL;
    uint64_t x = seed;
    x = mix64(x + G * (match + 1));
    x = mix64(x + G * (turn + 1));
    x = mix64(x + G * (kind + 1));
    x = mix64(x + G * (key + 1));
    return x;
}

/* Fills unit_actions[u_cap] / tile_actions[t_cap] of one match from its packed state. */
int lux_oracle_gen_actions(const LuxHeader* hdr, const LuxCell* cells, const LuxUnit* units,
                           const uint16_t* tiles, uint64_t seed, uint64_t match_id,
                           uint32_t* unit_actions, uint8_t* tile_actions, int u_cap, int t_cap) {
    if (!hdr || !cells || !units || !tiles || !unit_actions || !tile_actions) return LUX_E_ARG;
    memset(unit_actions, 0, sizeof(uint32_t) * u_cap);
    memset(tile_actions, 0, t_cap);
    if (hdr->done) return 0;
    uint64_t turn = (uint64_t)hdr->turn;
    for (int i = 0; i < hdr->n_units; i++) {
        const LuxUnit* u = &units[i];
        uint64_t r = lux_oracle_hash(seed, match_id, turn, 0, (uint64_t)u->id);
        int can_act = u->cd_q < 4;
        if (!can_act && ((r >> 32) & 15) != 0) continue;
        unsigned p = (unsigned)(r % 100);
        uint32_t w = 0;
        const LuxCell* uc = &cells[u->y * hdr->size + u->x];
        int cargo = u->wood + u->coal + u->uranium;
        int night = (hdr->turn % 40) >= 30;
        int can_build = !(u->team_type & 2) && cargo >= 100 && !is_tile(uc) && uc->amount == 0;
        if (can_build && ((r >> 40) & 3) != 0) w = LUX_UA_BCITY;
        else if (night && is_tile(uc) && ((r >> 42) & 7) != 0) w = 0;
        else if (p < 40) w = 0;
        else if (p < 80) w = LUX_UA_MOVE | ((1 + ((r >> 8) & 3)) << 3);
        else if (p < 85) w = LUX_UA_BCITY;
        else if (p < 88) w = LUX_UA_PILLAGE;
        else {
            int best = -1;
            for (int j = 0; j < hdr->n_units; j++) {
                const LuxUnit* v = &units[j];
                if (j == i || ((v->team_type ^ u->team_type) & 1)) continue;
                int d = abs((int)v->x - (int)u->x) + abs((int)v->y - (int)u->y);
                if (d <= 1 && (best < 0 || v->id < units[best].id)) best = j;
            }
            if (best >= 0) {
                uint32_t res = (uint32_t)((r >> 8) % 3), amount = 1 + (uint32_t)((r >> 16) % 100);
                w = LUX_UA_TRANSFER | (res << 3) | (amount << 6) | ((uint32_t)units[best].id << 17);
            }
        }
        unit_actions[i] = w;
    }
    for (int p = 0; p < hdr->n_tiles; p++) {
        const LuxCell* c = &cells[tiles[p]];
        uint64_t r = lux_oracle_hash(seed, match_id, turn, 1, (uint64_t)tiles[p]);
        int can_act = c->tile_cd < 1;
        if (!can_act && ((r >> 32) & 15) != 0) continue;
        unsigned q = (unsigned)(r % 100);
        tile_actions[p] = q < 45 ? LUX_TA_BUILD_WORKER : q < 50 ? LUX_TA_BUILD_CART : q < 90 ? LUX_TA_RESEARCH : LUX_TA_NONE;
    }
    return 0;
}

int lux_oracle_gen_actions_batch(const LuxBatch* b, uint64_t seed, uint64_t match_id_base,
                                 uint32_t* unit_actions, uint8_t* tile_actions, int first, int count) {
    if (!b) return LUX_E_ARG;
    int ncell = b->size * b->size;
    for (int m = first; m < first + count && m < b->n_matches; m++) {
        int rc = lux_oracle_gen_actions(&b->hdr[m], b->cells + (size_t)m * ncell, b->units + (size_t)m * b->u_cap,
                                        b->tiles + (size_t)m * b->t_cap, seed, match_id_base + (uint64_t)m,
                                        unit_actions + (size_t)m * b->u_cap, tile_actions + (size_t)m * b->t_cap,
                                        b->u_cap, b->t_cap);
        if (rc) return rc;
    }
    return 0;
}

/* ---------------------------------------------------------------- timed multi-turn driver (CPU baseline)
 * Runs `turns` turns of matches [first, first+count) with the canonical action stream entirely in C
 * (one call per host thread).  Only the step calls are timed; action construction is not.
 * Finished matches are replaced by pool entries exactly like lux_b200_reset_done
 * (hash of match index and epoch) when pool != NULL. */
#include <time.h>

static double now_sec(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int lux_oracle_run_batch(const LuxBatch* b, const LuxBatch* pool, uint64_t seed, uint64_t match_id_base,
                         uint32_t* unit_actions, uint8_t* tile_actions, int first, int count, int turns,
                         uint32_t epoch0, double* step_seconds, long long* live_turns) {
    if (!b || !unit_actions || !tile_actions || !step_seconds || !live_turns) return LUX_E_ARG;
    int ncell = b->size * b->size;
    double spent = 0.0;
    long long live = 0;
    for (int t = 0; t < turns; t++) {
        int rc = lux_oracle_gen_actions_batch(b, seed, match_id_base, unit_actions, tile_actions, first, count);
        if (rc) return rc;
        for (int m = first; m < first + count && m < b->n_matches; m++) live += !b->hdr[m].done;
        double t0 = now_sec();
        rc = lux_oracle_step_batch(b, unit_actions, tile_actions, first, count);
        spent += now_sec() - t0;
        if (rc) return rc;
        if (pool) {
            for (int m = first; m < first + count && m < b->n_matches; m++) {
                if (!b->hdr[m].done) continue;
                uint32_t hsh = (uint32_t)m * 0x9E3779B1u + (epoch0 + (uint32_t)t) * 0x85EBCA6Bu;
                hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du; hsh ^= hsh >> 12;
                int src = (int)(hsh % (uint32_t)pool->n_matches);
                memcpy(b->cells + (size_t)m * ncell, pool->cells + (size_t)src * ncell, (size_t)ncell * 8);
                memcpy(b->units + (size_t)m * b->u_cap, pool->units + (size_t)src * b->u_cap, (size_t)b->u_cap * 16);
                memcpy(b->cities + (size_t)m * b->c_cap, pool->cities + (size_t)src * b->c_cap, (size_t)b->c_cap * 16);
                memcpy(b->tiles + (size_t)m * b->t_cap, pool->tiles + (size_t)src * b->t_cap, (size_t)b->t_cap * 2);
                memcpy(b->res + (size_t)m * b->r_cap, pool->res + (size_t)src * b->r_cap, (size_t)b->r_cap * sizeof(LuxRes));
                b->hdr[m] = pool->hdr[src];
            }
        }
    }
    *step_seconds = spent;
    *live_turns = live;
    return 0;
}

/* ---------------------------------------------------------------- observations
 * Restatement of AgentPolicy.get_observation, /root/reference/examples/agent_policy.py:188-424, for
 * every entity MatchController would yield for `team` (match_controller.py:271-289): the team's
 * units with can_act() in dict order, then its city tiles with can_act() in cities x cells order.
 * Keeps the reference's structure: per-key node lists (:197-247), np.delete of the closest node for
 * the self-exclusion (:328-336), argmin/argmax of squared distance with first-index ties (:17-22),
 * greedy direction_to (position.py:48-66).  float64 arithmetic, stored as float32. */
#define OBS_DIM 85

typedef struct { int n; int x[MAXCELLS], y[MAXCELLS]; } NodeList;

static int obs_direction_slot(int px, int py, int tx, int ty) {
    /* position.py:48-66 checks N, E, S, W and keeps the first strict improvement;
       agent_policy.py:349-355 maps c,n,w,s,e -> 0,1,2,3,4 */
    int best = abs(tx - px) + abs(ty - py), dir = 0;
    const int cx[4] = {0, 1, 0, -1}, cy[4] = {-1, 0, 1, 0}, slot[4] = {1, 4, 3, 2};
    for (int k = 0; k < 4; k++) {
        int d = abs(tx - (px + cx[k])) + abs(ty - (py + cy[k]));
        if (d < best) { best = d; dir = slot[k]; }
    }
    return dir;
}

static int obs_arg(const NodeList* l, int px, int py, int furthest) {
    int best = -1; long bd = 0;
    for (int i = 0; i < l->n; i++) {
        long d = (long)(l->x[i] - px) * (l->x[i] - px) + (long)(l->y[i] - py) * (l->y[i] - py);
        if (best < 0 || (furthest ? d > bd : d < bd)) { best = i; bd = d; }
    }
    return best;
}

int lux_oracle_observe(const LuxHeader* h, const LuxCell* cells, const LuxUnit* units, const LuxCity* cities,
                       const uint16_t* tiles, const LuxRes* res, int team, float* obs, int16_t* ent_slot,
                       int32_t* ent_count, int e_cap) {
    if (!h || !cells || !units || !cities || !tiles || !res || !obs || !ent_slot || !ent_count) return LUX_E_ARG;
    *ent_count = 0;
    if (h->done) return 0;
    int S = h->size;
    static __thread NodeList nodes[5];      /* wood, coal, uranium, own city tiles, own workers */
    for (int k = 0; k < 5; k++) nodes[k].n = 0;
    for (int r = 0; r < h->n_res; r++) {     /* game.map.resources order; depleted cells are not in the list */
        const int rc = LUX_RES_CELL(res[r]);
        const LuxCell* c = &cells[rc];
        if (c->amount == 0) continue;
        NodeList* l = &nodes[res_type(c) - 1];
        l->x[l->n] = rc % S; l->y[l->n] = rc / S; l->n++;
    }
    int own_workers = 0, own_carts = 0, own_tiles = 0, opp_tiles = 0;
    for (int i = 0; i < h->n_units; i++) {
        const LuxUnit* u = &units[i];
        if ((u->team_type & 1) != team) continue;
        if (u->team_type & 2) { own_carts++; continue; }
        NodeList* l = &nodes[4];
        l->x[l->n] = u->x; l->y[l->n] = u->y; l->n++;
        own_workers++;
    }
    for (int p = 0; p < h->n_tiles; p++) {
        const LuxCell* c = &cells[tiles[p]];
        if (tile_team(c) != team) { opp_tiles++; continue; }
        NodeList* l = &nodes[3];
        l->x[l->n] = tiles[p] % S; l->y[l->n] = tiles[p] / S; l->n++;
        own_tiles++;
    }
    int night = is_night(h);
    int n_ent = 0;
    int total = h->n_units + h->n_tiles;
    for (int e = 0; e < total; e++) {
        int is_unit = e < h->n_units;
        int px, py, kind, slot_code, ucell_units = 0;
        const LuxUnit* U = NULL;
        if (is_unit) {
            U = &units[e];
            if ((U->team_type & 1) != team || !(U->cd_q < 4)) continue;
            px = U->x; py = U->y; kind = (U->team_type & 2) ? 1 : 0; slot_code = e;
            for (int i = 0; i < h->n_units; i++) ucell_units += (units[i].x == px && units[i].y == py);
        } else {
            int p = e - h->n_units;
            const LuxCell* c = &cells[tiles[p]];
            if (tile_team(c) != team || !(c->tile_cd < 1)) continue;
            px = tiles[p] % S; py = tiles[p] / S; kind = 2; slot_code = 0x4000 | p;
        }
        if (n_ent >= e_cap) return LUX_E_CAPACITY;
        float* o = obs + (size_t)n_ent * OBS_DIM;
        for (int k = 0; k < OBS_DIM; k++) o[k] = 0.0f;
        o[kind] = 1.0f;
        int at = 3;
        for (int furthest = 0; furthest < 2; furthest++) {
            for (int key = 0; key < 5; key++, at += 7) {
                const NodeList* full = &nodes[key];
                if (full->n == 0) continue;                               /* key not in object_nodes */
                static __thread NodeList filt;
                const NodeList* l = full;
                if ((key == 3 && kind == 2) || (kind == 0 && key == 4 && ucell_units <= 1)) {
                    int drop = obs_arg(full, px, py, 0);                  /* np.delete(closest) :333-334 */
                    filt.n = 0;
                    for (int i = 0; i < full->n; i++)
                        if (i != drop) { filt.x[filt.n] = full->x[i]; filt.y[filt.n] = full->y[i]; filt.n++; }
                    l = &filt;
                }
                if (l->n == 0) { o[at + 5] = 1.0f; continue; }
                int idx = obs_arg(l, px, py, furthest);
                int tx = l->x[idx], ty = l->y[idx];
                o[at + obs_direction_slot(px, py, tx, ty)] = 1.0f;
                double dist = (double)(abs(tx - px) + abs(ty - py)) / 20.0;
                o[at + 5] = (float)(dist < 1.0 ? dist : 1.0);
                const LuxCell* tc = &cells[ty * S + tx];
                double v;
                if (key == 3) {
                    const LuxCity* city = &cities[tc->city_slot];
                    double upkeep = (double)(city->ntiles * UPKEEP_CITY - city->adjsum * ADJ_BONUS);
                    v = (double)city->fuel / (upkeep * 200.0);
                } else if (key < 3) {
                    v = (double)tc->amount / 500.0;
                } else {
                    /* first unit in that cell's dict (:380-381); every unit standing on a cell gives the same
                       value in reachable states, the first in slot order is taken */
                    const LuxUnit* f = NULL;
                    for (int i = 0; i < h->n_units && !f; i++)
                        if (units[i].x == tx && units[i].y == ty) f = &units[i];
                    int space = ((f->team_type & 2) ? CAP_CART : CAP_WORKER) - (f->wood + f->coal + f->uranium);
                    v = (double)space / 100.0;
                }
                o[at + 6] = (float)(v < 1.0 ? v : 1.0);
            }
        }
        if (is_unit) {
            int space = ((U->team_type & 2) ? CAP_CART : CAP_WORKER) - (U->wood + U->coal + U->uranium);
            o[73] = (float)((double)space / 100.0);
        }
        o[74] = night ? 1.0f : 0.0f;
        o[75] = (float)((double)h->turn / 360.0);
        o[76] = (float)((double)own_tiles / 30.0);
        o[77] = (float)((double)opp_tiles / 30.0);
        o[78] = (float)((double)own_workers / 30.0);
        o[79] = (float)((double)own_workers / 30.0);   /* agent_policy.py:214-215 iterates the own team twice */
        o[80] = (float)((double)own_carts / 30.0);
        o[81] = (float)((double)own_carts / 30.0);
        o[82] = (float)((double)h->research[team] / 200.0);
        o[83] = h->research[team] >= RESEARCH_COAL ? 1.0f : 0.0f;
        o[84] = h->research[team] >= RESEARCH_URANIUM ? 1.0f : 0.0f;
        ent_slot[n_ent] = (int16_t)slot_code;
        n_ent++;
    }
    *ent_count = n_ent;
    return 0;
}

int lux_oracle_observe_batch(const LuxBatch* b, const uint8_t* team_sel, float* obs, int16_t* ent_slot,
                             int32_t* ent_count, int e_cap) {
    if (!b || !team_sel) return LUX_E_ARG;
    int ncell = b->size * b->size;
    for (int m = 0; m < b->n_matches; m++) {
        int rc = lux_oracle_observe(&b->hdr[m], b->cells + (size_t)m * ncell, b->units + (size_t)m * b->u_cap,
                                    b->cities + (size_t)m * b->c_cap, b->tiles + (size_t)m * b->t_cap,
                                    b->res + (size_t)m * b->r_cap, team_sel[m], obs + (size_t)m * e_cap * OBS_DIM,
                                    ent_slot + (size_t)m * e_cap, ent_count + m, e_cap);
        if (rc) return rc;
    }
    return 0;
}

/* ---------------------------------------------------------------- MatchController layer
 * lux_oracle_prevalidate: Action.is_valid as MatchController.take_action applies it
 * (match_controller.py:137-187) to each decision in yield order -- per team, units in dict order and
 * then city tiles -- against the actions already buffered this turn.  Invalid actions are zeroed in
 * the packed tensors; lux_oracle_step then sees what Game.run_turn_with_actions would be handed.
 * Spawn and build-city decisions need no filtering here: their is_valid (actions.py:159-194,
 * :256-285) equals validate_command (game.py:656-718) in the canonical order. */
int lux_oracle_prevalidate(const LuxHeader* h, const LuxCell* cells, const LuxUnit* units, const uint16_t* tiles,
                           uint32_t* ua, uint8_t* ta) {
    if (!h || !cells || !units || !tiles || !ua || !ta) return LUX_E_ARG;
    int S = h->size, n = h->n_units;
    static __thread int buffered_target[MAXU];      /* target cell of a buffered move, -1 none */
    for (int i = 0; i < n; i++) buffered_target[i] = -1;
    for (int i = 0; i < n; i++) {                    /* slot order == team A then team B, dict order */
        const LuxUnit* u = &units[i];
        int team = u->team_type & 1;
        int kind = ua[i] & 7;
        int can_act = u->cd_q < 4;
        if (kind == LUX_UA_MOVE) {                   /* MoveAction.is_valid actions.py:58-116 */
            int dir = (ua[i] >> 3) & 7, ok = can_act && dir <= 4;
            int nx = u->x, ny = u->y;
            if (ok) { nx += DX[dir]; ny += DY[dir]; ok = nx >= 0 && ny >= 0 && nx < S && ny < S; }
            if (ok) {
                int tgt = ny * S + nx;
                const LuxCell* tc = &cells[tgt];
                if (is_tile(tc) && tile_team(tc) != team) ok = 0;
                if (ok && !is_tile(tc)) {
                    int around[5], na = adjacent_cells(S, tgt, around);
                    around[na++] = tgt;
                    for (int k = 0; k < na && ok; k++)
                        for (int j = 0; j < n && ok; j++) {
                            const LuxUnit* v = &units[j];
                            if (j == i || (v->team_type & 1) != team || v->y * S + v->x != around[k]) continue;
                            if (buffered_target[j] == tgt) ok = 0;          /* :104-108 */
                            if (nx == u->x && ny == u->y) ok = 0;           /* :110-113 (compares with the mover) */
                        }
                }
                if (ok) buffered_target[i] = tgt;
            }
            if (!ok) ua[i] = 0;
        } else if (kind == LUX_UA_TRANSFER) {        /* TransferAction.is_valid actions.py:322-345 */
            int dest = (int)(ua[i] >> 17), ok = 0;
            for (int j = 0; j < n; j++) {
                const LuxUnit* v = &units[j];
                if ((v->team_type & 1) == team && v->id == dest) {
                    ok = j != i && can_act && abs((int)v->x - (int)u->x) + abs((int)v->y - (int)u->y) <= 1;
                    break;
                }
            }
            if (!ok) ua[i] = 0;
        } else if (kind == LUX_UA_PILLAGE) {         /* PillageAction.is_valid actions.py:368-385 */
            if (!can_act) ua[i] = 0;
        }
    }
    for (int p = 0; p < h->n_tiles; p++)             /* ResearchAction.is_valid actions.py:412-438 */
        if (ta[p] == LUX_TA_RESEARCH && !(cells[tiles[p]].tile_cd < 1)) ta[p] = 0;
    return 0;
}

/* AgentPolicy.action_code_to_action (examples/agent_policy.py:117-132, :426-470) for the entities of
 * one team in observation order: unit codes % 9 = center, n, w, s, e, transfer->cart, transfer->worker,
 * build city, pillage; tile codes % 3 = build worker, build cart, research.  Transfers follow
 * smart_transfer_to_nearby (:24-100). */
static int smart_transfer(const LuxHeader* h, const LuxUnit* units, int i, int want_cart, uint32_t* word) {
    const LuxUnit* u = &units[i];
    int S = h->size, team = u->team_type & 1;
    int amounts[3] = {u->wood, u->uranium, u->coal};          /* cargo dict order: wood, uranium, coal */
    const int codes[3] = {0, 2, 1};                           /* packed resource field: wood 0, coal 1, uranium 2 */
    int rtype = -1, ramount = 0;
    for (int k = 0; k < 3; k++) if (amounts[k] > ramount) { rtype = codes[k]; ramount = amounts[k]; }
    int adj[4], na = adjacent_cells(S, u->y * S + u->x, adj), target = -1;
    for (int k = 0; k < na; k++)
        for (int j = 0; j < h->n_units; j++) {                /* cell.units order: slot order stands in for arrival order */
            const LuxUnit* v = &units[j];
            if (v->y * S + v->x != adj[k] || ((v->team_type >> 1) & 1) != want_cart || (v->team_type & 1) != team) continue;
            if (target < 0) { target = j; continue; }
            int sv = (want_cart ? CAP_CART : CAP_WORKER) - (v->wood + v->coal + v->uranium);
            const LuxUnit* t = &units[target];
            int st = (want_cart ? CAP_CART : CAP_WORKER) - (t->wood + t->coal + t->uranium);
            if (sv >= ramount && st >= ramount) { if (sv < st) target = j; }
            else if (st >= ramount) { }
            else if (sv > st) target = j;
        }
    if (target < 0 || rtype < 0) return 0;                    /* TransferAction with None fields: is_valid False */
    const LuxUnit* t = &units[target];
    int st = (want_cart ? CAP_CART : CAP_WORKER) - (t->wood + t->coal + t->uranium);
    if (st < ramount) ramount = st;
    if (ramount > 2047) ramount = 2047;
    *word = LUX_UA_TRANSFER | ((uint32_t)rtype << 3) | ((uint32_t)ramount << 6) | ((uint32_t)t->id << 17);
    return 1;
}

int lux_oracle_encode_actions(const LuxHeader* h, const LuxUnit* units, const int16_t* ent_slot, int ent_count,
                              const int32_t* codes, uint32_t* ua, uint8_t* ta, int u_cap, int t_cap) {
    if (!h || !units || !ent_slot || !codes || !ua || !ta) return LUX_E_ARG;
    memset(ua, 0, sizeof(uint32_t) * u_cap);
    memset(ta, 0, t_cap);
    if (h->done) return 0;
    static const int dirs[5] = {LUX_DIR_CENTER, LUX_DIR_NORTH, LUX_DIR_WEST, LUX_DIR_SOUTH, LUX_DIR_EAST};
    for (int e = 0; e < ent_count; e++) {
        int code = codes[e], slot = ent_slot[e];
        if (slot & 0x4000) { ta[slot & 0x3FFF] = (uint8_t)(1 + ((code % 3) + 3) % 3); continue; }
        int c = ((code % 9) + 9) % 9;
        uint32_t w = 0;
        if (c < 5) w = LUX_UA_MOVE | (dirs[c] << 3);
        else if (c == 5) smart_transfer(h, units, slot, 1, &w);
        else if (c == 6) smart_transfer(h, units, slot, 0, &w);
        else if (c == 7) w = LUX_UA_BCITY;
        else w = LUX_UA_PILLAGE;
        ua[slot] = w;
    }
    return 0;
}

/* AgentPolicy.get_reward (examples/agent_policy.py:494-560) evaluated once per turn for `team`:
 * last[3] = units_last, city_tiles_last, fuel_collected_last (game_start :480-492 zeroes them). */
int lux_oracle_reward(const LuxHeader* h, int team, int32_t* last, double* reward) {
    if (!h || !last || !reward) return LUX_E_ARG;
    int units_now = h->n_team_units[team], tiles_now = h->n_team_tiles[team], fuel = h->fuel_generated[team];
    double r = 0.0;
    r += (double)(units_now - last[0]) * 0.05;
    r += (double)(tiles_now - last[1]) * 0.1;
    r += (double)(fuel - last[2]) / 20000.0;
    r += h->done ? (double)tiles_now : 0.0;
    last[0] = units_now; last[1] = tiles_now; last[2] = fuel;
    *reward = r;
    return 0;
}
