// lux_step.cu -- sm_100a kernels and the C ABI (include/lux_b200.h) of the batched turn engine.
//
// lux_step_kernel: one warp per match.  The warp's slice of dynamic shared memory holds the
// whole match; it is filled by the TMA unit (cp.async.bulk global->shared, completion on a
// per-warp mbarrier: 128-byte header first, then the live prefixes of every section), the
// turn runs in shared memory (lux_step_core.cuh), and the cell grid + header go back with
// cp.async.bulk shared->global while the compacted unit/city/tile lists are written with
// plain vector stores.  No CPU fallback exists: without a device the entry points return
// LUX_E_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "lux_step_core.cuh"
#include "lux_tma.cuh"

#define LUX_B200_VERSION 100

// ---------------------------------------------------------------- the step kernel
// Speculative first batch: these many leading records of each list travel together with the header
// and the cell grid, so a typical (sparse) match needs ONE global->shared round trip; only matches
// with more entities issue a second batch for the remainder.
#define LUX_PRE_UNITS 32
#define LUX_PRE_CITIES 16
#define LUX_PRE_TILES 32

// Launch modes.  The staged footprint of a match is dominated by capacity, not by content: with the
// default 32x32 capacities (252 units / 255 cities / 384 tiles) a match needs 25 KB of shared memory and
// only 8 warps fit on an SM, although a typical match holds a dozen entities.  So a step is two launches:
//   kSmall : every match; shared memory is sized for `stage` capacities (e.g. 64/64/96); a match whose
//            populations could outgrow them during this turn is appended to a retry list instead;
//   kList  : the retry list, full capacities, grid-stride over the list.
// kAll is the single full-capacity launch (small maps, dense batches, the non-TMA path).
enum { kAll = 0, kSmall = 1, kList = 2 };

struct LuxStage { int u, c, t; };      // shared-memory staging capacities (units, cities, tiles)

// result of one match: skipped because finished, did not fit the staging class, or played
// (LUX_M_HEAVY: its end-of-turn population is above the "schedule first next step" threshold)
enum { LUX_M_DONE = 0, LUX_M_SPILL = 1, LUX_M_PLAYED = 2, LUX_M_HEAVY = 3 };
#define LUX_HEAVY_THRESHOLD 28     // units + tiles

// One lane group plays match `m` (or, with valid == 0, walks along on an empty match: the groups of a
// warp run in lockstep and every collective needs all 32 lanes).
template <bool kTma, int kS, int G>
__device__ __forceinline__ int lux_step_match(const LuxBatch& b, int m, int valid, const uint32_t* __restrict__ unit_actions,
                                               const uint8_t* __restrict__ tile_actions, const LuxSmemLayout& L,
                                               unsigned char* smem, unsigned flags, const LuxStage& stage,
                                               bool check_fit, uint32_t& phase) {
    const int lane = lux_lane<G>();
    const int size = kS > 0 ? kS : b.size;
    const int ncell = size * size;
    if (!valid) m = 0;
    LuxHeader* g_hdr = b.hdr + m;
    LuxCell* g_cells = b.cells + (size_t)m * ncell;
    LuxUnit* g_units = b.units + (size_t)m * b.u_cap;
    LuxCity* g_cities = b.cities + (size_t)m * b.c_cap;
    uint16_t* g_tiles = b.tiles + (size_t)m * b.t_cap;
    const uint16_t* g_res = b.res + (size_t)m * b.r_cap;
    const uint32_t* g_ua = unit_actions + (size_t)m * b.u_cap;
    const uint8_t* g_ta = tile_actions + (size_t)m * b.t_cap;
    LuxMatch M;
    lux_match_bind(M, smem, L, g_cells, size, b.u_cap, b.c_cap, b.t_cap, b.r_cap);
    int active = valid, result = LUX_M_DONE;

    if (kTma) {
        uint64_t* bar = (uint64_t*)(smem + L.mbar);
        const uint32_t pu = min(stage.u, LUX_PRE_UNITS), pc = min(stage.c, LUX_PRE_CITIES), pt = min(stage.t, LUX_PRE_TILES);
        const uint32_t b_res = b.r_cap * 2;
        if (active && lane == 0) {
            mbar_expect_tx(bar, LUX_HDR_BYTES + b_res + pu * 16 + pc * 16 + pt * 2 + pu * 4 + pt);
            tma_g2s(M.hdr, g_hdr, LUX_HDR_BYTES, bar);
            tma_g2s(M.units, g_units, pu * 16, bar);
            tma_g2s(M.uact, g_ua, pu * 4, bar);
            tma_g2s(M.cities, g_cities, pc * 16, bar);
            tma_g2s(M.tiles, g_tiles, pt * 2, bar);
            tma_g2s(M.tact, g_ta, pt, bar);
            tma_g2s(M.res, g_res, b_res, bar);
        }
        lux_sync<G>();
        if (active) { mbar_wait(bar, phase); phase ^= 1; }
        uint32_t nu = 0, nc = 0, nt = 0;
        if (active) {
            if (M.hdr->done) active = 0;
            nu = M.hdr->n_units; nc = M.hdr->n_cities; nt = M.hdr->n_tiles;
        }
        // the live lists themselves must fit the staging arrays before anything more is loaded
        if (active && check_fit && (nu > (uint32_t)stage.u || nc > (uint32_t)stage.c || nt > (uint32_t)stage.t)) {
            active = 0; result = LUX_M_SPILL;
        }
        const int more = active && (nu > pu || nc > pc || nt > pt);   // group-uniform
        if (more && lane == 0) {
            const uint32_t ru = nu > pu ? nu - pu : 0, rc = nc > pc ? nc - pc : 0;
            const uint32_t rt2 = nt > pt ? ((nt - pt) * 2 + 15) & ~15u : 0;
            const uint32_t rua = nu > pu ? ((nu - pu) * 4 + 15) & ~15u : 0, rta = nt > pt ? ((nt - pt) + 15) & ~15u : 0;
            mbar_expect_tx(bar, ru * 16 + rc * 16 + rt2 + rua + rta);
            if (ru) tma_g2s(M.units + pu, g_units + pu, ru * 16, bar);
            if (rc) tma_g2s(M.cities + pc, g_cities + pc, rc * 16, bar);
            if (rt2) tma_g2s(M.tiles + pt, g_tiles + pt, rt2, bar);
            if (rua) tma_g2s(M.uact + pu, g_ua + pu, rua, bar);
            if (rta) tma_g2s(M.tact + pt, g_ta + pt, rta, bar);
        }
        lux_sync<G>();        // also: every lane has read the header before an idle group clears it
        if (more) { mbar_wait(bar, phase); phase ^= 1; }
    } else {
        if (active) lux_copy16<G>(M.hdr, g_hdr, LUX_HDR_BYTES);
        lux_sync<G>();
        if (active && M.hdr->done) active = 0;
        lux_sync<G>();
        if (active) {
            lux_copy16<G>(M.units, g_units, M.hdr->n_units * 16);
            lux_copy16<G>(M.cities, g_cities, M.hdr->n_cities * 16);
            lux_copy16<G>(M.tiles, g_tiles, M.hdr->n_tiles * 2);
            lux_copy16<G>(M.res, g_res, M.hdr->n_res * 2);
            lux_copy16<G>(M.uact, g_ua, M.hdr->n_units * 4);
            lux_copy16<G>(M.tact, g_ta, M.hdr->n_tiles);
        }
        lux_sync<G>();
    }

    if (check_fit) {
        // growth inside this turn is bounded by the actions actually present: every spawn action may add
        // a unit, every build-city action a tile and a city
        int spawns = 0, builds = 0;
        const int nu = active ? M.hdr->n_units : 0, nt = active ? M.hdr->n_tiles : 0, nc = active ? M.hdr->n_cities : 0;
        const int nt_max = lux_wmax<G>(nt), nu_max = lux_wmax<G>(nu);
        for (int base = 0; base < nt_max; base += G) {
            int p = base + lane;
            int code = p < nt ? M.tact[p] : 0;
            spawns += __popc(lux_ballot<G>(code == LUX_TA_BUILD_WORKER || code == LUX_TA_BUILD_CART));
        }
        for (int base = 0; base < nu_max; base += G) {
            int u = base + lane;
            int kind = u < nu ? (int)(M.uact[u] & 7) : 0;
            builds += __popc(lux_ballot<G>(kind == LUX_UA_BCITY));
        }
        if (active && (nu + spawns > stage.u || nc + builds > stage.c || nt + builds > stage.t)) { active = 0; result = LUX_M_SPILL; }
    }
    if (!active) lux_idle_header<G>(M);
    lux_sync<G>();

    LuxTurnOut T = lux_turn<kS, G>(M, flags);
    lux_finish_and_store<kS, G>(M, T, g_hdr, g_units, g_cities, g_tiles, active);
    if (!active) return result;
    return (!M.hdr->done && (int)M.hdr->n_units + (int)M.hdr->n_tiles > LUX_HEAVY_THRESHOLD) ? LUX_M_HEAVY : LUX_M_PLAYED;
}

// Scheduling scratch (device pointers, owned by the library, see StepScratch below).
//  * retry list: matches the small staging class could not hold this step (consumed by kList);
//  * heavy list: matches whose population was above LUX_HEAVY_THRESHOLD at the end of the previous step.
//    The first `hcap` CTAs of a launch take their match from this list, the remaining CTAs walk the
//    matches in index order and skip the ones flagged in `hint_cur`: long jobs start first, which removes
//    the tail a late-starting dense match would otherwise add to the launch (LPT scheduling).
//    Lists and hints are hints only: a stale or empty list just means index order.
struct LuxSched {
    int* retry_cnt; int* retry_list;                 // this step
    const int* heavy_cnt; const int* heavy_list; const uint8_t* hint_cur;   // built during the previous step
    int* heavy_next_cnt; int* heavy_next; uint8_t* hint_next;               // built now
    int* zero_a; int* zero_b;                        // counters to clear for the step after
    int hcap;                                        // 0 = scheduling off
};

template <bool kTma, int kS, int kMode, int G>
__global__ void __launch_bounds__((G == 32 && kMode != kList) ? 1024 : 64, 1)
lux_step_kernel(LuxBatch b, const uint32_t* __restrict__ unit_actions, const uint8_t* __restrict__ tile_actions,
                LuxSmemLayout L, unsigned flags, LuxStage stage, LuxSched sc) {
    extern __shared__ __align__(128) unsigned char lux_smem[];
    // a group of G lanes plays one match; 32 / G matches share a warp
    const int grp = threadIdx.x / G, lane = lux_lane<G>();
    const int gpc = blockDim.x / G;
    unsigned char* smem = lux_smem + (size_t)grp * L.total;
    uint32_t phase = 0;
    if (kTma) {
        if (lane == 0) {
            mbar_init((uint64_t*)(smem + L.mbar), 1);
            fence_mbar_init();
        }
    }
    {   // the per-cell byte planes start clean; every turn leaves them clean again
        const int ncell = (kS > 0 ? kS : b.size) * (kS > 0 ? kS : b.size);
        lux_zero16<G>(smem + L.p1, (ncell + 15) & ~15);
        lux_zero16<G>(smem + L.occ, (ncell + 15) & ~15);
    }
    lux_sync<G>();
    if (kMode == kList) {
        const int count = *sc.retry_cnt;
        for (int idx = blockIdx.x * gpc + grp; lux_wany<G>(idx < count); idx += gridDim.x * gpc) {
            const int valid = idx < count;
            lux_step_match<kTma, kS, G>(b, valid ? sc.retry_list[idx] : 0, valid, unit_actions, tile_actions, L, smem, flags, stage,
                                        false, phase);
            lux_sync<G>();
        }
        return;
    }
    const int gidx = blockIdx.x * gpc + grp;
    int m = 0, valid = 1;
    if (sc.zero_a && gidx == 0 && lane == 0) { *sc.zero_a = 0; *sc.zero_b = 0; }
    if (sc.hcap > 0) {
        if (gidx < sc.hcap) {
            int cnt = *sc.heavy_cnt;
            if (gidx >= (cnt < sc.hcap ? cnt : sc.hcap)) valid = 0;
            else m = sc.heavy_list[gidx];
        } else {
            m = gidx - sc.hcap;
            if (m >= b.n_matches || sc.hint_cur[m]) valid = 0;
        }
    } else {
        m = gidx;
    }
    if (m < 0 || m >= b.n_matches) valid = 0;
    if (!lux_wany<G>(valid)) return;
    int r = lux_step_match<kTma, kS, G>(b, m, valid, unit_actions, tile_actions, L, smem, flags, stage, kMode == kSmall, phase);
    if (valid && lane == 0) {
        if (r == LUX_M_SPILL) sc.retry_list[atomicAdd(sc.retry_cnt, 1)] = m;
        if (sc.hcap > 0) {
            uint8_t heavy = 0;
            if (r == LUX_M_SPILL || r == LUX_M_HEAVY) {
                int i = atomicAdd(sc.heavy_next_cnt, 1);
                if (i < sc.hcap) { sc.heavy_next[i] = m; heavy = 1; }
            }
            sc.hint_next[m] = heavy;
        }
    }
}

// kernels this library has launched (all entry points of this translation unit); bench.py reports the
// difference over its timed region as "gpu_launches"
static unsigned long long g_launches = 0;
static int g_carveout = -1;       // preferred shared-memory carve-out in percent (-1 = driver default)
extern "C" unsigned long long lux_b200_launch_count(void) { return g_launches; }

// `gpc` = matches (lane groups) per CTA; the CTA has gpc * G threads
template <bool kTma, int kS, int kMode, int G>
static int launch_step(const LuxBatch* b, const uint32_t* ua, const uint8_t* ta, const LuxSmemLayout& L, unsigned flags,
                       LuxStage stage, const LuxSched& sc, int n_groups, int wpc, cudaStream_t st) {
    const int gpc = wpc * (32 / G);
    size_t smem = (size_t)L.total * gpc;
    cudaError_t e = cudaFuncSetAttribute(lux_step_kernel<kTma, kS, kMode, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return LUX_E_CUDA;
    if (g_carveout >= 0) {
        // the grid is read in place through L1: leave part of the SM's 256 KB to the cache
        e = cudaFuncSetAttribute(lux_step_kernel<kTma, kS, kMode, G>, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout);
        if (e != cudaSuccess) return LUX_E_CUDA;
    }
    lux_step_kernel<kTma, kS, kMode, G><<<(n_groups + gpc - 1) / gpc, 32 * wpc, smem, st>>>(*b, ua, ta, L, flags, stage, sc);
    g_launches++;
    return LUX_OK;
}

template <bool kTma, int kMode, int G>
static int launch_step_size(const LuxBatch* b, const uint32_t* ua, const uint8_t* ta, const LuxSmemLayout& L, unsigned flags,
                            LuxStage stage, const LuxSched& sc, int n_groups, int wpc, cudaStream_t st) {
    switch (kTma ? b->size : 0) {
        case 12: return launch_step<kTma, 12, kMode, G>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
        case 16: return launch_step<kTma, 16, kMode, G>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
        case 24: return launch_step<kTma, 24, kMode, G>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
        case 32: return launch_step<kTma, 32, kMode, G>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
        default: return launch_step<kTma, 0, kMode, G>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
    }
}

// lane-group width: 8, 16 or 32 lanes per match (kList, the dense class, always plays with a full warp)
template <bool kTma, int kMode>
static int launch_step_group(int group, const LuxBatch* b, const uint32_t* ua, const uint8_t* ta, const LuxSmemLayout& L,
                             unsigned flags, LuxStage stage, const LuxSched& sc, int n_groups, int wpc, cudaStream_t st) {
    if constexpr (kTma && kMode != kList) {
        if (group == 8) return launch_step_size<kTma, kMode, 8>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
        if (group == 16) return launch_step_size<kTma, kMode, 16>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
    }
    return launch_step_size<kTma, kMode, 32>(b, ua, ta, L, flags, stage, sc, n_groups, wpc, st);
}

// ---------------------------------------------------------------- canonical action stream
__device__ __forceinline__ uint64_t lux_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__device__ __forceinline__ uint64_t lux_hash(uint64_t seed, uint64_t match, uint64_t turn, uint64_t kind, uint64_t key) {
    const uint64_t G = 0x9E3779B97F4A7C15ULL;
    uint64_t x = seed;
    x = lux_mix64(x + G * (match + 1));
    x = lux_mix64(x + G * (turn + 1));
    x = lux_mix64(x + G * (kind + 1));
    x = lux_mix64(x + G * (key + 1));
    return x;
}

// r % m for a 64-bit r with 32-bit arithmetic only (2^32 = q*m + c): exact, and far cheaper than the
// 64-bit remainder routine
__device__ __forceinline__ unsigned lux_mod64(uint64_t r, unsigned m) {
    unsigned hi = (unsigned)(r >> 32), lo = (unsigned)r;
    unsigned c = (unsigned)(4294967296ULL % m);
    return ((hi % m) * c + lo % m) % m;
}

// one warp per match, state read in place from global memory (harness kernel, not the hot path)
__device__ __forceinline__ void lux_gen_actions_match(const LuxBatch& b, int m, uint64_t seed, uint64_t match_id_base,
                                                      uint32_t* __restrict__ unit_actions,
                                                      uint8_t* __restrict__ tile_actions, unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    const int ncell = b.size * b.size, S = b.size;
    const LuxHeader* h = b.hdr + m;
    const LuxCell* cells = b.cells + (size_t)m * ncell;
    const LuxUnit* units = b.units + (size_t)m * b.u_cap;
    const uint16_t* tiles = b.tiles + (size_t)m * b.t_cap;
    uint32_t* ua = unit_actions + (size_t)m * b.u_cap;
    uint8_t* ta = tile_actions + (size_t)m * b.t_cap;
    const int nu = h->n_units, nt = h->n_tiles, turn = h->turn, done = h->done;
    if (counters && lane == 0 && !done) {
        // live match-turns and their algorithmic bytes for the turn about to run
        // (`counters` is the CTA's shared-memory pair here; one global add per CTA follows)
        atomicAdd(counters, 1ULL);
        atomicAdd(counters + 1, (unsigned long long)(2 * (8 * ncell + 16 * nu + 16 * (int)h->n_cities + 64) + 4 * (nu + nt)));
    }
    const uint64_t mid = match_id_base + (uint64_t)m;
    // only the live prefix (rounded up to the staging granule) is written: the step kernel never looks further
    const int nu_w = done ? 0 : min(b.u_cap, max((nu + 3) & ~3, LUX_PRE_UNITS));
    const int nt_w = done ? 0 : min(b.t_cap, max((nt + 15) & ~15, LUX_PRE_TILES));
    for (int i = lane; i < nu_w; i += 32) {
        uint32_t w = 0;
        if (!done && i < nu) {
            const LuxUnit U = units[i];
            uint64_t r = lux_hash(seed, mid, (uint64_t)turn, 0, (uint64_t)U.id);
            int can_act = U.cd_q < 4;
            if (can_act || ((r >> 32) & 15) == 0) {
                unsigned p = lux_mod64(r, 100);
                const LuxCell C = cells[U.y * S + U.x];
                int cargo = U.wood + U.coal + U.uranium;
                int night = (turn % 40) >= 30;
                int on_tile = (C.flags & 0x0C) != 0;
                int can_build = !(U.team_type & 2) && cargo >= 100 && !on_tile && C.amount == 0;
                if (can_build && ((r >> 40) & 3) != 0) w = LUX_UA_BCITY;
                else if (night && on_tile && ((r >> 42) & 7) != 0) w = 0;
                else if (p < 40) w = 0;
                else if (p < 80) w = LUX_UA_MOVE | ((1 + ((r >> 8) & 3)) << 3);
                else if (p < 85) w = LUX_UA_BCITY;
                else if (p < 88) w = LUX_UA_PILLAGE;
                else {
                    int best_id = 0x7fffffff;
                    for (int j = 0; j < nu; j++) {
                        const LuxUnit V = units[j];
                        if (j == i || ((V.team_type ^ U.team_type) & 1)) continue;
                        int d = abs((int)V.x - (int)U.x) + abs((int)V.y - (int)U.y);
                        if (d <= 1 && V.id < best_id) best_id = V.id;
                    }
                    if (best_id != 0x7fffffff) {
                        uint32_t res = lux_mod64(r >> 8, 3), amount = 1 + lux_mod64(r >> 16, 100);
                        w = LUX_UA_TRANSFER | (res << 3) | (amount << 6) | ((uint32_t)best_id << 17);
                    }
                }
            }
        }
        ua[i] = w;
    }
    for (int p = lane; p < nt_w; p += 32) {
        uint8_t code = 0;
        if (!done && p < nt) {
            int cell = tiles[p];
            uint64_t r = lux_hash(seed, mid, (uint64_t)turn, 1, (uint64_t)cell);
            int can_act = cells[cell].tile_cd < 1;
            if (can_act || ((r >> 32) & 15) == 0) {
                unsigned q = lux_mod64(r, 100);
                code = q < 45 ? LUX_TA_BUILD_WORKER : q < 50 ? LUX_TA_BUILD_CART : q < 90 ? LUX_TA_RESEARCH : LUX_TA_NONE;
            }
        }
        ta[p] = code;
    }
}

__global__ void lux_gen_actions_kernel(LuxBatch b, uint64_t seed, uint64_t match_id_base,
                                       uint32_t* __restrict__ unit_actions, uint8_t* __restrict__ tile_actions,
                                       unsigned long long* counters) {
    __shared__ unsigned long long cta_counters[2];
    if (threadIdx.x < 2) cta_counters[threadIdx.x] = 0;
    __syncthreads();
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m < b.n_matches) lux_gen_actions_match(b, m, seed, match_id_base, unit_actions, tile_actions, counters ? cta_counters : nullptr);
    __syncthreads();
    if (counters && threadIdx.x < 2 && cta_counters[threadIdx.x]) atomicAdd(counters + threadIdx.x, cta_counters[threadIdx.x]);
}

// ---------------------------------------------------------------- episode statistics
__global__ void lux_stats_kernel(LuxBatch b, unsigned long long* out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    long long v[12];
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = 0;
    if (m < b.n_matches) {
        const LuxHeader* h = b.hdr + m;
        int done = h->done;
        v[0] = !done; v[1] = done;
        v[2] = done && h->winner == LUX_WIN_A; v[3] = done && h->winner == LUX_WIN_B; v[4] = done && h->winner == LUX_WIN_TIE;
        v[5] = h->turn;
        v[6] = (long long)h->fuel_generated[0] + h->fuel_generated[1];
        v[7] = h->n_units; v[8] = h->n_tiles; v[9] = h->n_cities;
        v[10] = h->status != 0;
        // algorithmic bytes of one turn: 2*(8*S^2 + 16*U + 16*K + 64) + 4*(U+T)
        v[11] = done ? 0 : 2LL * (8LL * b.size * b.size + 16LL * h->n_units + 16LL * h->n_cities + 64) + 4LL * (h->n_units + h->n_tiles);
    }
#pragma unroll
    for (int k = 0; k < 12; k++) {
        long long x = v[k];
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(out + k, (unsigned long long)x);
    }
}

// ---------------------------------------------------------------- reset finished matches from a pool
__global__ void lux_reset_done_kernel(LuxBatch b, LuxBatch pool, uint32_t epoch, int* reset_count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= b.n_matches) return;
    if (!b.hdr[m].done) return;
    int src = 0;
    if (lane == 0) {
        // deterministic choice (replayable trajectories): a hash of (match, epoch)
        uint32_t hsh = (uint32_t)m * 0x9E3779B1u + epoch * 0x85EBCA6Bu;
        hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du; hsh ^= hsh >> 12;
        src = (int)(hsh % (uint32_t)pool.n_matches);
        if (reset_count) atomicAdd(reset_count, 1);
    }
    src = __shfl_sync(0xffffffffu, src, 0);
    const int ncell = b.size * b.size;
    // 8 independent 16-byte loads in flight per lane before the stores: the copy of one match is a
    // single warp's latency chain otherwise (measured 17 us for 9 KB)
    auto copy = [&](void* d, const void* s, size_t bytes) {
        const uint4* sp = (const uint4*)s;
        uint4* dp = (uint4*)d;
        const size_t n16 = bytes / 16;
        for (size_t base = 0; base < n16; base += 32 * 8) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                size_t i = base + (size_t)k * 32 + lane;
                if (i < n16) v[k] = sp[i];
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                size_t i = base + (size_t)k * 32 + lane;
                if (i < n16) dp[i] = v[k];
            }
        }
    };
    copy(b.cells + (size_t)m * ncell, pool.cells + (size_t)src * ncell, (size_t)ncell * 8);
    copy(b.units + (size_t)m * b.u_cap, pool.units + (size_t)src * b.u_cap, (size_t)pool.hdr[src].n_units * 16);
    copy(b.cities + (size_t)m * b.c_cap, pool.cities + (size_t)src * b.c_cap, (size_t)pool.hdr[src].n_cities * 16);
    copy(b.tiles + (size_t)m * b.t_cap, pool.tiles + (size_t)src * b.t_cap, ((size_t)pool.hdr[src].n_tiles * 2 + 15) & ~(size_t)15);
    copy(b.res + (size_t)m * b.r_cap, pool.res + (size_t)src * b.r_cap, ((size_t)pool.hdr[src].n_res * 2 + 15) & ~(size_t)15);
    __syncwarp();
    copy(b.hdr + m, pool.hdr + src, LUX_HDR_BYTES);
}

// ---------------------------------------------------------------- C ABI
static int check_batch(const LuxBatch* b) {
    if (!b) return LUX_E_ARG;
    if (b->n_matches < 0 || b->size < 4 || b->size > 32 || (b->size & 1)) return LUX_E_ARG;
    if (b->u_cap < 4 || b->u_cap > 252 || (b->u_cap & 3)) return LUX_E_CAPACITY;
    if (b->c_cap < 1 || b->c_cap > 255) return LUX_E_CAPACITY;
    if (b->t_cap < 16 || b->t_cap > 1024 || (b->t_cap & 15)) return LUX_E_CAPACITY;
    if (b->r_cap < 8 || b->r_cap > 1024 || (b->r_cap & 7)) return LUX_E_CAPACITY;
    if (!b->hdr || !b->cells || !b->units || !b->cities || !b->tiles || !b->res) return LUX_E_ARG;
    if (((uintptr_t)b->hdr | (uintptr_t)b->cells | (uintptr_t)b->units | (uintptr_t)b->cities | (uintptr_t)b->tiles |
         (uintptr_t)b->res) & 15)
        return LUX_E_ARG;
    return LUX_OK;
}

static int cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return LUX_OK;
    fprintf(stderr, "lux_b200: %s: %s\n", what, cudaGetErrorString(e));
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? LUX_E_NO_DEVICE : LUX_E_CUDA;
}

extern "C" int lux_b200_version(void) { return LUX_B200_VERSION; }

extern "C" int lux_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" size_t lux_b200_step_smem_bytes(int size, int u_cap, int c_cap, int t_cap, int r_cap) {
    return (size_t)lux_smem_layout(size, u_cap, c_cap, t_cap, r_cap).total;
}

static int g_use_tma = 1;
static int g_warps_per_cta = 0;   // 0 = automatic
static int g_two_class = -1;      // -1 automatic, 0 off, 1 always
static int g_stage[3] = {96, 64, 96};   // small staging class (units, cities, tiles); tests shrink it to force spills
static int g_group = 0;           // lanes per match: 0 automatic, else 8 / 16 / 32
static int g_heavy_first = 0;     // schedule the previous step's populous matches first (measured +-0: off)
// tuning / fallback knobs (tests flip them to cover every launch path)
extern "C" int lux_b200_set_option(const char* name, int value) {
    if (!name) return LUX_E_ARG;
    if (!strcmp(name, "tma")) { g_use_tma = value != 0; return LUX_OK; }
    if (!strcmp(name, "warps_per_cta")) { g_warps_per_cta = value; return LUX_OK; }
    if (!strcmp(name, "two_class")) { g_two_class = value; return LUX_OK; }
    if (!strcmp(name, "heavy_first")) { g_heavy_first = value != 0; return LUX_OK; }
    if (!strcmp(name, "group")) { g_group = value; return LUX_OK; }
    if (!strcmp(name, "carveout")) { g_carveout = value > 100 ? 100 : value; return LUX_OK; }
    if (!strcmp(name, "stage_units")) { g_stage[0] = value < 32 ? 32 : (value + 3) & ~3; return LUX_OK; }
    if (!strcmp(name, "stage_cities")) { g_stage[1] = value < 16 ? 16 : value; return LUX_OK; }
    if (!strcmp(name, "stage_tiles")) { g_stage[2] = value < 32 ? 32 : (value + 15) & ~15; return LUX_OK; }
    return LUX_E_ARG;
}

// Scheduling scratch of one batch (keyed by its header pointer; a few batches are cached so that
// interleaved batches -- mixed map sizes, a pool -- each keep their own lists).
struct StepScratch {
    const void* key; int n; int dev;
    int* counts;                 // [0..2] heavy counts (3-way rotation), [3..5] retry counts
    int* lists;                  // heavy[2][n] then retry[n]
    unsigned char* hints;        // [2][n]
    int* seen;                   // pinned host copy of a recent retry count (may be stale)
    uint2* entries[2];           // sparse host actions of this / the previous call (device copies)
    int entries_cap, entries_prev_n, entries_flip;
    uint8_t* done_dev;           // compact done flags for the host path
    int* summary_dev;            // [0] max n_units, [1] max n_tiles, [2] live matches, [3] finished
    unsigned step;
    int dense_hold;              // steps left in single-launch mode after a dense observation
};
static StepScratch g_scratch[8];
static unsigned g_scratch_clock = 0;

static void scratch_free(StepScratch& s) {
    if (s.counts) cudaFree(s.counts);
    if (s.lists) cudaFree(s.lists);
    if (s.hints) cudaFree(s.hints);
    if (s.seen) cudaFreeHost(s.seen);
    for (int k = 0; k < 2; k++) if (s.entries[k]) cudaFree(s.entries[k]);
    if (s.done_dev) cudaFree(s.done_dev);
    if (s.summary_dev) cudaFree(s.summary_dev);
    memset(&s, 0, sizeof s);
}

static StepScratch* scratch_get(const LuxBatch* b, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    StepScratch* slot = nullptr;
    for (auto& s : g_scratch)
        if (s.key == (const void*)b->hdr && s.n == b->n_matches && s.dev == dev) return &s;
    for (auto& s : g_scratch)
        if (!s.key) { slot = &s; break; }
    if (!slot) { slot = &g_scratch[g_scratch_clock++ % 8]; cudaStreamSynchronize(st); scratch_free(*slot); }
    size_t n = (size_t)b->n_matches;
    if (cudaMalloc(&slot->counts, 8 * sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMalloc(&slot->lists, 3 * n * sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMalloc(&slot->hints, 2 * n) != cudaSuccess) return nullptr;
    if (cudaMallocHost(&slot->seen, sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMemsetAsync(slot->counts, 0, 8 * sizeof(int), st) != cudaSuccess) return nullptr;
    if (cudaMemsetAsync(slot->hints, 0, 2 * n, st) != cudaSuccess) return nullptr;
    *slot->seen = 0;
    slot->key = b->hdr; slot->n = b->n_matches; slot->dev = dev; slot->step = 0; slot->dense_hold = 0;
    return slot;
}

extern "C" int lux_b200_step(const LuxBatch* b, const uint32_t* unit_actions, const uint8_t* tile_actions,
                             uint32_t flags, void* cuda_stream) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!unit_actions || !tile_actions) return LUX_E_ARG;
    if (((uintptr_t)unit_actions | (uintptr_t)tile_actions) & 15) return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    LuxSmemLayout L = lux_smem_layout(b->size, b->u_cap, b->c_cap, b->t_cap, b->r_cap);
    LuxStage full = {b->u_cap, b->c_cap, b->t_cap};
    if ((size_t)L.total > 227 * 1024) return LUX_E_CAPACITY;
    // CTA shape: `wpc` warps, each holding 32 / group matches.  Automatic: one warp per CTA unless the
    // per-CTA footprint is so small that the 32-CTA-per-SM limit would bind.
    auto pick_wpc = [&](const LuxSmemLayout& lay, int group) {
        size_t per_warp = (size_t)lay.total * (32 / group);
        int w = g_warps_per_cta > 0 ? g_warps_per_cta : (per_warp >= 7 * 1024 ? 1 : 2);
        while (w > 1 && per_warp * w > 227 * 1024) w--;
        return w;
    };
    auto pick_group = [&](const LuxSmemLayout& lay) {
        int g = g_group == 8 || g_group == 16 || g_group == 32 ? g_group : 32;   // automatic: measured fastest
        while (g < 32 && (size_t)lay.total * (32 / g) > 227 * 1024) g *= 2;
        return g;
    };
    LuxSched sc;
    memset(&sc, 0, sizeof sc);
    if (!g_use_tma) {
        int wpc = pick_wpc(L, 32);
        rc = launch_step_group<false, kAll>(32, b, unit_actions, tile_actions, L, flags, full, sc, b->n_matches, wpc, st);
        if (rc) return rc;
        return cuda_ok(cudaGetLastError(), "lux_step_kernel launch");
    }

    LuxStage small = {g_stage[0] < b->u_cap ? g_stage[0] : b->u_cap, g_stage[1] < b->c_cap ? g_stage[1] : b->c_cap,
                      g_stage[2] < b->t_cap ? g_stage[2] : b->t_cap};
    LuxSmemLayout Ls = lux_smem_layout(b->size, small.u, small.c, small.t, b->r_cap);
    bool two = g_two_class != 0 && (g_two_class == 1 || (L.total >= 10 * 1024 && Ls.total * 4 <= L.total * 3));
    bool sched = g_heavy_first != 0;
    StepScratch* S = (two || sched) ? scratch_get(b, st) : nullptr;
    if ((two || sched) && !S) return LUX_E_CUDA;
    if (two && g_two_class < 0) {
        if (S->dense_hold > 0) { S->dense_hold--; two = false; }
        else if (*S->seen * 2 > b->n_matches) { S->dense_hold = 64; *S->seen = 0; two = false; }
    }
    int hcap = 0;
    if (S) {
        const size_t n = (size_t)b->n_matches;
        const unsigned t = S->step++;
        sc.retry_cnt = S->counts + 3 + t % 3;
        sc.retry_list = S->lists + 2 * n;
        sc.heavy_cnt = S->counts + t % 3;
        sc.heavy_list = S->lists + (t & 1) * n;
        sc.hint_cur = S->hints + (t & 1) * n;
        sc.heavy_next_cnt = S->counts + (t + 1) % 3;
        sc.heavy_next = S->lists + ((t + 1) & 1) * n;
        sc.hint_next = S->hints + ((t + 1) & 1) * n;
        sc.zero_a = S->counts + (t + 2) % 3;
        sc.zero_b = S->counts + 3 + (t + 1) % 3;
        hcap = sched ? (b->n_matches / 4 + 64) : 0;
        sc.hcap = hcap;
    }
    if (!two) {
        // single full-capacity launch (small maps, dense batches); spills cannot happen
        int group = g_group == 8 || g_group == 16 ? pick_group(L) : 32;
        rc = launch_step_group<true, kAll>(group, b, unit_actions, tile_actions, L, flags, full, sc, b->n_matches + hcap,
                                           pick_wpc(L, group), st);
        if (rc) return rc;
        return cuda_ok(cudaGetLastError(), "lux_step_kernel launch");
    }
    int group = pick_group(Ls);
    rc = launch_step_group<true, kSmall>(group, b, unit_actions, tile_actions, Ls, flags, small, sc, b->n_matches + hcap,
                                         pick_wpc(Ls, group), st);
    if (rc) return rc;
    rc = cuda_ok(cudaGetLastError(), "lux_step_kernel (small class) launch");
    if (rc) return rc;
    int grid_l = b->n_matches < 148 * 8 ? b->n_matches : 148 * 8;
    rc = launch_step_group<true, kList>(32, b, unit_actions, tile_actions, L, flags, full, sc, grid_l, 1, st);
    if (rc) return rc;
    rc = cuda_ok(cudaGetLastError(), "lux_step_kernel (retry list) launch");
    if (rc) return rc;
    // trail the retry count to the host without synchronising; a stale value only delays the mode switch
    return cuda_ok(cudaMemcpyAsync(S->seen, sc.retry_cnt, sizeof(int), cudaMemcpyDeviceToHost, st), "retry count copy");
}

extern "C" int lux_b200_step_host(const LuxBatch* b, const uint32_t* unit_actions_host,
                                  const uint8_t* tile_actions_host, uint32_t* unit_actions_dev,
                                  uint8_t* tile_actions_dev, uint8_t* done_host, LuxHeader* hdr_host,
                                  int u_used, int t_used, uint32_t flags, void* cuda_stream) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!unit_actions_host || !tile_actions_host || !unit_actions_dev || !tile_actions_dev) return LUX_E_ARG;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    size_t n = (size_t)b->n_matches;
    // only the first u_used / t_used slots of every match travel (0 = all); the caller knows the largest
    // live counts from the headers of the previous step
    if (u_used <= 0 || u_used > b->u_cap) u_used = b->u_cap;
    if (t_used <= 0 || t_used > b->t_cap) t_used = b->t_cap;
    u_used = (u_used + 3) & ~3;
    t_used = (t_used + 15) & ~15;
    if (u_used > b->u_cap) u_used = b->u_cap;
    if (t_used > b->t_cap) t_used = b->t_cap;
    rc = cuda_ok(cudaMemcpy2DAsync(unit_actions_dev, (size_t)b->u_cap * 4, unit_actions_host, (size_t)b->u_cap * 4,
                                   (size_t)u_used * 4, n, cudaMemcpyHostToDevice, st), "h2d unit actions");
    if (rc) return rc;
    rc = cuda_ok(cudaMemcpy2DAsync(tile_actions_dev, (size_t)b->t_cap, tile_actions_host, (size_t)b->t_cap,
                                   (size_t)t_used, n, cudaMemcpyHostToDevice, st), "h2d tile actions");
    if (rc) return rc;
    rc = lux_b200_step(b, unit_actions_dev, tile_actions_dev, flags, cuda_stream);
    if (rc) return rc;
    if (hdr_host) {
        rc = cuda_ok(cudaMemcpyAsync(hdr_host, b->hdr, n * LUX_HDR_BYTES, cudaMemcpyDeviceToHost, st), "d2h headers");
        if (rc) return rc;
    }
    rc = cuda_ok(cudaStreamSynchronize(st), "step sync");
    if (rc) return rc;
    if (done_host && hdr_host)
        for (size_t i = 0; i < n; i++) done_host[i] = hdr_host[i].done;
    return LUX_OK;
}

// ---------------------------------------------------------------- sparse host actions
// entry.x = tile flag (bit 31) | match index (bits 30:10) | slot (bits 9:0), entry.y = action word / byte
__global__ void lux_apply_entries_kernel(const uint2* __restrict__ entries, int k, int clear, uint32_t* __restrict__ ua,
                                         uint8_t* __restrict__ ta, int u_cap, int t_cap, int n_matches) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    uint2 e = entries[i];
    unsigned m = (e.x >> 10) & 0x1FFFFFu, slot = e.x & 1023u;
    if ((int)m >= n_matches) return;
    if (e.x >> 31) { if ((int)slot < t_cap) ta[(size_t)m * t_cap + slot] = clear ? 0 : (uint8_t)e.y; }
    else if ((int)slot < u_cap) ua[(size_t)m * u_cap + slot] = clear ? 0u : e.y;
}

__global__ void lux_summary_kernel(LuxBatch b, uint8_t* __restrict__ done, int* __restrict__ summary) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int nu = 0, nt = 0, live = 0, fin = 0;
    if (m < b.n_matches) {
        const LuxHeader* h = b.hdr + m;
        int d = h->done;
        done[m] = (uint8_t)d;
        nu = h->n_units; nt = h->n_tiles; live = !d; fin = d;
    }
    nu = __reduce_max_sync(0xffffffffu, nu);
    nt = __reduce_max_sync(0xffffffffu, nt);
    live = __reduce_add_sync(0xffffffffu, live);
    fin = __reduce_add_sync(0xffffffffu, fin);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(summary, nu); atomicMax(summary + 1, nt);
        if (live) atomicAdd(summary + 2, live);
        if (fin) atomicAdd(summary + 3, fin);
    }
}

// The turn for a HOST caller that holds an action LIST (what Game.run_turn_with_actions takes): k entries
// are uploaded, scattered into the device action tensors (the previous call's entries are cleared first),
// the step runs, and done flags (n bytes) + a 4-int summary come back.  Per step this moves 8*k bytes up
// and n + 16 bytes down instead of dense action tensors and full headers.
extern "C" int lux_b200_step_host_sparse(const LuxBatch* b, const void* entries_host, int k, uint32_t* unit_actions_dev,
                                         uint8_t* tile_actions_dev, uint8_t* done_host, int32_t* summary_host,
                                         uint32_t flags, void* cuda_stream) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (k < 0 || (k > 0 && !entries_host) || !unit_actions_dev || !tile_actions_dev) return LUX_E_ARG;
    if (b->n_matches >= (1 << 21)) return LUX_E_ARG;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    StepScratch* S = scratch_get(b, st);
    if (!S) return LUX_E_CUDA;
    if (!S->done_dev) {
        if (cudaMalloc(&S->done_dev, (size_t)b->n_matches) != cudaSuccess) return LUX_E_CUDA;
        if (cudaMalloc(&S->summary_dev, 4 * sizeof(int)) != cudaSuccess) return LUX_E_CUDA;
    }
    if (k > S->entries_cap) {
        cudaStreamSynchronize(st);
        int cap = k * 2 + 1024;
        for (int j = 0; j < 2; j++) {
            uint2* fresh = nullptr;
            if (cudaMalloc(&fresh, (size_t)cap * sizeof(uint2)) != cudaSuccess) return LUX_E_CUDA;
            if (S->entries[j]) {
                cudaMemcpy(fresh, S->entries[j], (size_t)S->entries_cap * sizeof(uint2), cudaMemcpyDeviceToDevice);
                cudaFree(S->entries[j]);
            }
            S->entries[j] = fresh;
        }
        S->entries_cap = cap;
    }
    uint2* prev = S->entries[S->entries_flip];
    uint2* cur = S->entries[S->entries_flip ^ 1];
    if (S->entries_prev_n > 0) {
        lux_apply_entries_kernel<<<(S->entries_prev_n + 255) / 256, 256, 0, st>>>(prev, S->entries_prev_n, 1, unit_actions_dev,
                                                                                  tile_actions_dev, b->u_cap, b->t_cap, b->n_matches);
        g_launches++;
    }
    if (k > 0) {
        rc = cuda_ok(cudaMemcpyAsync(cur, entries_host, (size_t)k * sizeof(uint2), cudaMemcpyHostToDevice, st), "h2d action entries");
        if (rc) return rc;
        lux_apply_entries_kernel<<<(k + 255) / 256, 256, 0, st>>>(cur, k, 0, unit_actions_dev, tile_actions_dev, b->u_cap,
                                                                  b->t_cap, b->n_matches);
        g_launches++;
    }
    S->entries_flip ^= 1;
    S->entries_prev_n = k;
    rc = lux_b200_step(b, unit_actions_dev, tile_actions_dev, flags, cuda_stream);
    if (rc) return rc;
    rc = cuda_ok(cudaMemsetAsync(S->summary_dev, 0, 4 * sizeof(int), st), "summary memset");
    if (rc) return rc;
    lux_summary_kernel<<<(b->n_matches + 255) / 256, 256, 0, st>>>(*b, S->done_dev, S->summary_dev);
    g_launches++;
    if (done_host) {
        rc = cuda_ok(cudaMemcpyAsync(done_host, S->done_dev, (size_t)b->n_matches, cudaMemcpyDeviceToHost, st), "d2h done flags");
        if (rc) return rc;
    }
    if (summary_host) {
        rc = cuda_ok(cudaMemcpyAsync(summary_host, S->summary_dev, 4 * sizeof(int), cudaMemcpyDeviceToHost, st), "d2h summary");
        if (rc) return rc;
    }
    return cuda_ok(cudaStreamSynchronize(st), "step sync");
}

extern "C" int lux_b200_gen_actions(const LuxBatch* b, uint64_t seed, uint64_t match_id_base,
                                    uint32_t* unit_actions, uint8_t* tile_actions, uint64_t* counters,
                                    void* cuda_stream) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!unit_actions || !tile_actions) return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    const int wpc = 8;
    lux_gen_actions_kernel<<<(b->n_matches + wpc - 1) / wpc, 32 * wpc, 0, (cudaStream_t)cuda_stream>>>(
        *b, seed, match_id_base, unit_actions, tile_actions, (unsigned long long*)counters);
    g_launches++;
    return cuda_ok(cudaGetLastError(), "lux_gen_actions_kernel launch");
}

extern "C" int lux_b200_reset_done(const LuxBatch* b, const LuxBatch* pool, uint32_t epoch,
                                   int32_t* reset_count, void* cuda_stream) {
    int rc = check_batch(b);
    if (rc) return rc;
    rc = check_batch(pool);
    if (rc) return rc;
    if (pool->n_matches < 1) return LUX_E_ARG;
    if (pool->size != b->size || pool->u_cap != b->u_cap || pool->c_cap != b->c_cap || pool->t_cap != b->t_cap ||
        pool->r_cap != b->r_cap)
        return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    const int wpc = 4;
    lux_reset_done_kernel<<<(b->n_matches + wpc - 1) / wpc, 32 * wpc, 0, (cudaStream_t)cuda_stream>>>(
        *b, *pool, epoch, reset_count);
    g_launches++;
    return cuda_ok(cudaGetLastError(), "lux_reset_done_kernel launch");
}

extern "C" int lux_b200_stats(const LuxBatch* b, int64_t* out16, void* cuda_stream) {
    int rc = check_batch(b);
    if (rc) return rc;
    if (!out16) return LUX_E_ARG;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    rc = cuda_ok(cudaMemsetAsync(out16, 0, 16 * sizeof(int64_t), st), "stats memset");
    if (rc) return rc;
    if (b->n_matches == 0) return LUX_OK;
    lux_stats_kernel<<<(b->n_matches + 127) / 128, 128, 0, st>>>(*b, (unsigned long long*)out16);
    g_launches++;
    return cuda_ok(cudaGetLastError(), "lux_stats_kernel launch");
}
