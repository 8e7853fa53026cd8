/*
 * lux_b200.h -- C ABI of the B200-native batched Lux AI 2021 turn engine.
 *
 * The reference (glmcdona/LuxPythonEnvGym) has no FFI layer: the per-turn path is the pure
 * Python call `Game.run_turn_with_actions(actions)` (luxai2021/game/game.py:390-535).  The
 * entry points below are what a binding for that path would call; each one names the
 * reference interface it replaces.  Plain pointers and sizes only, no torch types; every
 * function returns 0 on success or a negative LUX_E_* code, never throws.
 *
 * All state pointers are DEVICE pointers unless the function name ends in `_host`.
 * A batch is `n_matches` independent matches of one map side `size`, stored as six
 * sections with a leading match dimension (see luxpythonenvgym_b200/state.py):
 *
 *   hdr    [n][128]            LuxHeader
 *   cells  [n][size*size]      LuxCell   (8 B)
 *   units  [n][u_cap]          LuxUnit   (16 B) team-A units in dict order, then team B
 *   cities [n][c_cap]          LuxCity   (16 B) Game.cities dict order
 *   tiles  [n][t_cap]          uint16    cell index of each city tile, cities x city_cells order
 *   res    [n][r_cap]          LuxRes    (4 B) every resource cell of the map in GameMap.resources order:
 *                                        cell index, resource type and a MIRROR of the cell's current amount
 */
#ifndef LUX_B200_H
#define LUX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUX_HDR_BYTES 128

/* resource codes, LuxCell.flags[1:0] */
enum { LUX_RES_NONE = 0, LUX_RES_WOOD = 1, LUX_RES_COAL = 2, LUX_RES_URANIUM = 3 };
/* unit action word, bits [2:0] (luxai2021/game/actions.py:44-438 action classes) */
enum { LUX_UA_NONE = 0, LUX_UA_MOVE = 1, LUX_UA_BCITY = 2, LUX_UA_PILLAGE = 3, LUX_UA_TRANSFER = 4 };
/* move direction, bits [5:3] of a move word (position.py:36-46) */
enum { LUX_DIR_CENTER = 0, LUX_DIR_NORTH = 1, LUX_DIR_EAST = 2, LUX_DIR_SOUTH = 3, LUX_DIR_WEST = 4 };
/* transfer word: [4:3] resource-1, [16:6] amount (0..2047), [31:17] destination unit id */
/* city-tile action byte */
enum { LUX_TA_NONE = 0, LUX_TA_BUILD_WORKER = 1, LUX_TA_BUILD_CART = 2, LUX_TA_RESEARCH = 3 };
/* LuxHeader.status bits */
enum { LUX_ST_UNIT_OVERFLOW = 1, LUX_ST_CITY_OVERFLOW = 2, LUX_ST_TILE_OVERFLOW = 4, LUX_ST_BAD_STATE = 8 };
/* LuxHeader.winner */
enum { LUX_WIN_A = 0, LUX_WIN_B = 1, LUX_WIN_TIE = 2, LUX_WIN_NONE = 255 };
/* lux_b200_step flags */
enum { LUX_STEP_PREVALIDATE = 1 };
/* error codes */
enum { LUX_OK = 0, LUX_E_ARG = -1, LUX_E_CAPACITY = -2, LUX_E_CUDA = -3, LUX_E_NO_DEVICE = -4 };

/* Every field is naturally aligned; the aligned attribute lets device code move a record with
 * one 8/16-byte access.  Section base addresses must be 16-byte aligned. */
#define LUX_ALIGNED(n) __attribute__((aligned(n)))
typedef struct LUX_ALIGNED(16) LuxHeader {            /* game.py:85-147 */
    int32_t  turn;
    int32_t  unit_id_count;            /* Game.global_unit_id_count */
    int32_t  city_id_count;            /* Game.global_city_id_count */
    uint16_t n_units, n_cities, n_tiles, n_res;
    uint16_t n_team_units[2];
    uint8_t  done;                     /* run_turn_with_actions returned True */
    uint8_t  winner;                   /* LUX_WIN_* once done */
    uint8_t  status;                   /* LUX_ST_* */
    uint8_t  size;                     /* map side */
    int32_t  research[2];              /* teamStates[t].researchPoints */
    int32_t  fuel_generated[2];        /* stats.teamStats[t].fuelGenerated */
    int32_t  collected[2][3];          /* resourcesCollected wood/coal/uranium */
    int32_t  workers_built[2];
    int32_t  carts_built[2];
    int32_t  roads_built_q[2];         /* roadsBuilt * 4 */
    int32_t  city_tiles_built[2];
    uint16_t n_team_tiles[2];          /* city tiles per team (derived from cities[]) */
    uint8_t  pad[24];
} LuxHeader;

typedef struct LUX_ALIGNED(8) LuxCell {              /* cell.py:20-33, city.py:53-66 */
    uint16_t amount;                   /* Resource.amount, 0 = none */
    uint8_t  flags;                    /* [1:0] resource type, [3:2] tile team+1, [6:4] adjacent_city_tiles */
    uint8_t  road_q;                   /* Cell.road * 4 */
    uint8_t  tile_cd;                  /* CityTile.cooldown */
    uint8_t  city_slot;                /* index into cities[] */
    uint16_t key;                      /* index in City.city_cells */
} LuxCell;

/* One entry of the resource list (game_map.py:53-58 `GameMap.resources`, static after reset: depleted cells stay
 * listed with amount 0).  `amount` mirrors cells[cell].amount -- every writer of one writes the other -- so that
 * the per-turn passes over all resource cells (depletion + wood regrowth game.py:498-510, :1081-1102; the
 * observation's resource nodes agent_policy.py:197-247) read one contiguous list instead of one 32-byte DRAM
 * sector per cell.  `rtype` is the type the cell was created with (it stays after depletion). */
typedef uint32_t LuxRes;               /* [9:0] cell index, [11:10] resource type (LUX_RES_*), [31:16] amount */
#define LUX_RES_CELL(e) ((int)((e) & 0x3FFu))
#define LUX_RES_TYPE(e) ((int)(((e) >> 10) & 3u))
#define LUX_RES_AMOUNT(e) ((int)((e) >> 16))
#define LUX_RES_MAKE(cell, rtype, amount) ((LuxRes)((cell) | ((rtype) << 10) | ((uint32_t)(amount) << 16)))

typedef struct LUX_ALIGNED(16) LuxUnit {              /* unit.py:15-35 */
    int32_t  id;                       /* N of "u_N" */
    uint8_t  x, y;
    uint8_t  team_type;                /* bit0 team, bit1 cart */
    uint8_t  cd_q;                     /* cooldown * 4 */
    uint16_t wood, coal, uranium;
    uint16_t pad;
} LuxUnit;

typedef struct LUX_ALIGNED(16) LuxCity {              /* city.py:15-50 */
    int32_t  id;                       /* N of "c_N" */
    int32_t  fuel;
    uint16_t ntiles;
    uint16_t adjsum;                   /* sum of adjacent_city_tiles over the city's tiles */
    uint8_t  team;
    uint8_t  pad[3];
} LuxCity;

#ifdef __cplusplus
static_assert(sizeof(LuxHeader) == 128 && sizeof(LuxCell) == 8 && sizeof(LuxUnit) == 16 && sizeof(LuxCity) == 16,
              "packed state layout");
#else
_Static_assert(sizeof(LuxHeader) == 128 && sizeof(LuxCell) == 8 && sizeof(LuxUnit) == 16 && sizeof(LuxCity) == 16,
               "packed state layout");
#endif

typedef struct LuxBatch {
    int32_t n_matches;
    int32_t size;                      /* map side (12/16/24/32; any 4..32 accepted) */
    int32_t u_cap, c_cap, t_cap, r_cap;
    LuxHeader* hdr;
    LuxCell*   cells;
    LuxUnit*   units;
    LuxCity*   cities;
    uint16_t*  tiles;
    LuxRes*    res;
} LuxBatch;

/* Library / build identification; also reports whether a CUDA device is usable. */
int lux_b200_version(void);
int lux_b200_device_count(void);

/* Tuning knobs: "tma" (1 = cp.async.bulk staging, 0 = plain vector loads/stores),
 * "warps_per_cta" (0 = automatic), "two_class" (-1 automatic, 0 off, 1 always), "fork" (1 = the big class runs as an
 * independent grid on a high-priority side stream, default; 0 = as the small class's programmatic primary),
 * "overlap" (1 = programmatic dependent launches along the chain), "big_share" (the small slice is the smallest
 * candidate that at most 1 / big_share of the batch exceeded; default 3), "mid_slice" (shared-memory bytes per match
 * of the big class: 0 = automatic, 12 KB, with the rare match beyond it played by a launch of its own with
 * full-capacity slices; -1 = full capacity for the whole big class), "over_own" (that launch on a stream of its own:
 * -1 = when the statistics have seen such matches, otherwise it follows the big class on its stream; 0 never; 1 always),
 * "cta_sync" (warps per CTA of the small class with a CTA-wide rendezvous at some phase boundaries of the turn, so that
 * the warps fetch a phase's code together: -1 = automatic, 8 on dense batches and none on sparse ones; 0 = never; n = always n;
 * 32 x 32 maps, flags 0 only), "sync_mask" (which of the nine rendezvous points are armed, default 0x54),
 * "order" (1 = the small class plays its heavy matches first, through order lists the classification pass builds, so that
 * the grid ends on light ones; default), "n_ord" / "ord_step" (tuning probes: force the number of need classes, 0 =
 * automatic = two on sparse batches, four on dense ones, and their spacing in percent of the slice),
 * "slice" (shared-memory bytes per match of the small class, 0 = default), "group" (lanes per match: 0 automatic,
 * 8 / 16 / 32), "carveout" (preferred shared-memory carve-out in percent, -1 = automatic), "list_grid" (CTAs of
 * the big-class launch, 0 = automatic), "gen_wpc" (matches per CTA of the action-stream kernel); measurement aids:
 * "trace" (see lux_b200_debug_trace), "only_class" (1 / 2: launch only the big / small class -- results incomplete).
 * The options are process-global and not synchronised: set them before the first step and from one thread
 * (the per-batch scratch the library caches is behind a lock; one host thread per GPU is the supported use).
 * Returns LUX_E_ARG for an unknown name. */
int lux_b200_set_option(const char* name, int value);

/*
 * Measurement aid (not part of the reference's interface): after lux_b200_set_option("trace", 1) every launch of
 * a two-class step stamps its first start and last end (%globaltimer, ns) into a device buffer; this call
 * synchronises the device and copies 40 words out: stamps [0,1] big-class launch, [4,5] small-class launch, [6,7]
 * classification pass (~0 / 0 where a launch did not run), then two histograms in 8 us bins from the start of the
 * classification pass: [8..23] small-class warps starting, [24..39] big-class warps ending.  scripts/trace_step.py prints the timeline.
 */
int lux_b200_debug_trace(unsigned long long* out8);

/* Number of kernels the step / gen_actions / reset_done / stats entry points have launched so far. */
unsigned long long lux_b200_launch_count(void);

/* Dynamic shared memory one match needs in lux_b200_step for these capacities. */
size_t lux_b200_step_smem_bytes(int size, int u_cap, int c_cap, int t_cap, int r_cap);

/*
 * One turn for every live match of the batch.
 * Replaces: Game.run_turn_with_actions(actions) -> bool, luxai2021/game/game.py:390-535
 * (phases: validate_command :646-723, handle_movement_actions :1104-1195, CityTile.turn
 * city.py:96-120, Worker/Cart.turn unit.py:162-269, distribute_all_resources :868-1001,
 * handle_resource_deposit :1003-1023, handle_night :537-558, regenerate_trees :1081-1102,
 * match_over :570-592, run_cooldowns :560-568), applied to the action list
 *   [tile_actions[m][p] for p in canonical tile order] + [unit_actions[m][i] for i in slot order].
 * unit_actions: [n][u_cap] uint32, tile_actions: [n][t_cap] uint8 (device).  Matches whose
 * hdr.done is set are left untouched.  The bool return value lands in hdr.done, the
 * winner (get_winning_team :594-634, coin flip reported as LUX_WIN_TIE) in hdr.winner.
 * flags: LUX_STEP_PREVALIDATE first applies MatchController.take_action's Action.is_valid filter
 * (luxai2021/game/match_controller.py:137-187, actions.py:58-438) to the action tensors: of the
 * same-team moves into one non-city cell only the first in unit order survives, transfers need an
 * acting source and an adjacent destination on the team, pillage / research need can_act().
 */
int lux_b200_step(const LuxBatch* b, const uint32_t* unit_actions, const uint8_t* tile_actions,
                  uint32_t flags, void* cuda_stream);

/*
 * Same call with HOST action buffers and HOST result buffers (pinned or pageable):
 * copies the actions to the device staging buffers the caller provides, steps, and
 * copies back done[n] (uint8) and the headers (n*128 B) if the pointers are non-NULL.
 * Only the first u_used unit slots / t_used tile slots of every match are uploaded (0 = all of them);
 * they must cover the live entities, e.g. the maxima of hdr.n_units / hdr.n_tiles seen in the headers
 * the previous call returned (a freshly reset match holds 2 + 2).
 * Synchronises the stream before returning.
 */
int lux_b200_step_host(const LuxBatch* b, const uint32_t* unit_actions_host,
                       const uint8_t* tile_actions_host, uint32_t* unit_actions_dev,
                       uint8_t* tile_actions_dev, uint8_t* done_host, LuxHeader* hdr_host,
                       int u_used, int t_used, uint32_t flags, void* cuda_stream);

/*
 * The turn for a host caller that holds an action LIST, as Game.run_turn_with_actions(actions) takes it
 * (game.py:390): k entries of two uint32 { tile flag (bit 31) | match index (bits 30:10) | slot (bits 9:0),
 * action word (units) or action byte (tiles) }.  The entries are uploaded and scattered into the device
 * action tensors -- which must be all-zero before the first call and must not be written by anything else
 * between calls: the step zeroes the action words of every match it plays, and entries that name a finished
 * match or a slot beyond the live lists are dropped, so the tensors are clean again after every call --
 * the step runs, then done flags (n_matches bytes) and a 4-int summary { max n_units, max n_tiles, live
 * matches, finished matches } arrive in the host buffers (either may be NULL; they travel through mapped pinned
 * memory) and the stream is synchronised.  8*k bytes travel up and n + 16 bytes down per turn.
 */
int lux_b200_step_host_sparse(const LuxBatch* b, const void* entries_host, int k, uint32_t* unit_actions_dev,
                              uint8_t* tile_actions_dev, uint8_t* done_host, int32_t* summary_host,
                              uint32_t flags, void* cuda_stream);

/*
 * The same turn in two halves, for a caller that has more work to queue behind the step before it blocks
 * (pool resets, the observation pass): _submit uploads the list and queues scatter, step and the done-flag /
 * summary pass without waiting; _wait blocks until that pass has published its results (it polls the sequence
 * number the pass writes behind them in mapped host memory, and falls back to synchronising the stream), then
 * fills the host buffers exactly as lux_b200_step_host_sparse does.  Work queued on the stream after _submit may
 * still be running when _wait returns.  lux_b200_step_host_sparse == _submit + _wait + a stream synchronisation.
 * _wait without a preceding _submit for the batch returns LUX_E_ARG.
 */
int lux_b200_step_host_sparse_submit(const LuxBatch* b, const void* entries_host, int k, uint32_t* unit_actions_dev,
                                     uint8_t* tile_actions_dev, uint32_t flags, void* cuda_stream);
int lux_b200_step_host_wait(const LuxBatch* b, uint8_t* done_host, int32_t* summary_host, void* cuda_stream);

/*
 * Canonical seeded action stream, written as packed action
 * tensors for every live match: action = f(hash64(seed, match_id_base+m, turn, kind, key)).
 * Harness utility (benchmarks, parity tests); a policy normally fills the tensors itself.
 * counters (device uint64[2], may be NULL): [0] += live matches, [1] += algorithmic bytes of
 * the turn about to run, 2*(8*S^2 + 16*U + 16*K + 64) + 4*(U+T) per live match.
 */
int lux_b200_gen_actions(const LuxBatch* b, uint64_t seed, uint64_t match_id_base,
                         uint32_t* unit_actions, uint8_t* tile_actions, uint64_t* counters,
                         void* cuda_stream);

/*
 * Per-entity observation tensors for team `team_sel[m]` of every match.
 * Replaces: AgentPolicy.get_observation, examples/agent_policy.py:188-424, called once
 * per actionable entity by LuxEnvironment.step (luxai2021/env/lux_env.py:146).  Entity
 * order is MatchController's yield order (match_controller.py:271-289): the team's units
 * with can_act() in dict order, then its city tiles with can_act() in cities x cells order.
 * obs: [n][e_cap][85] float32, ent_slot: [n][e_cap] int16 (unit slot, or 0x4000|tile
 * position), ent_count: [n] int32.
 * Packed layout (row_offsets != NULL, int64[n+1], device): the rows of match m start at row
 * row_offsets[m] of obs[R][85] / ent_slot[R] and may use up to row_offsets[m+1] - row_offsets[m] rows (an
 * upper bound such as the team's units + tiles); unused rows of that range get ent_slot = -1.
 */
int lux_b200_observe(const LuxBatch* b, const uint8_t* team_sel, float* obs, int16_t* ent_slot,
                     int32_t* ent_count, int e_cap, const int64_t* row_offsets, void* cuda_stream);

/*
 * AgentPolicy action codes -> packed action tensors for the entities lux_b200_observe listed.
 * Replaces: AgentPolicy.action_code_to_action, examples/agent_policy.py:426-470 with the action
 * tables :117-132 (unit code % 9: center, n, w, s, e, transfer to a nearby cart, transfer to a nearby
 * worker (smart_transfer_to_nearby :24-100), build city, pillage; tile code % 3: build worker, build
 * cart, research).  codes: [n][e_cap] int32, or [R] with the packed row_offsets of lux_b200_observe.
 * Slots without an entity get "no action".
 */
int lux_b200_encode_actions(const LuxBatch* b, const int16_t* ent_slot, const int32_t* ent_count,
                            const int32_t* codes, int e_cap, const int64_t* row_offsets, uint32_t* unit_actions,
                            uint8_t* tile_actions, void* cuda_stream);

/*
 * Per-turn reward of team `team_sel[m]`.
 * Replaces: AgentPolicy.get_reward, examples/agent_policy.py:494-560 evaluated at a turn boundary:
 * 0.05*d(units) + 0.1*d(city tiles) + d(fuelGenerated)/20000 (+ city tiles when the match just ended).
 * reward_state: [n][4] int32 in/out (units_last, city_tiles_last, fuel_collected_last, unused; zero it
 * at match start like game_start :480-492), reward: [n] float64.
 */
int lux_b200_reward(const LuxBatch* b, const uint8_t* team_sel, int32_t* reward_state, double* reward,
                    void* cuda_stream);

/*
 * Replace every finished match (hdr.done != 0) by a fresh one copied from a pool of
 * initial states (same size/capacities); the pool entry is a hash of (match index, epoch)
 * so that a run is replayable.
 * Replaces: Game.reset(), luxai2021/game/game.py:76-157 (map generation itself stays on
 * the host: luxpythonenvgym_b200/mapgen.py).  reset_count (device int32[1], may be NULL) += resets.
 */
int lux_b200_reset_done(const LuxBatch* b, const LuxBatch* pool, uint32_t epoch,
                        int32_t* reset_count, void* cuda_stream);
/* The same with the epoch read from DEVICE memory (*epoch_dev) when the kernel runs: a caller that replays one
 * captured CUDA graph per turn advances the counter on the device. */
int lux_b200_reset_done_dev(const LuxBatch* b, const LuxBatch* pool, const uint32_t* epoch_dev,
                            int32_t* reset_count, void* cuda_stream);


/*
 * lux_b200_reset_done followed by lux_b200_gen_actions in one pass over the batch (the per-turn harness loop
 * of benchmarks and stress tests: one launch and one read of the headers instead of two).  Same arguments and
 * the same results as the two calls in that order.
 */
int lux_b200_reset_gen(const LuxBatch* b, const LuxBatch* pool, uint32_t epoch, int32_t* reset_count, uint64_t seed,
                       uint64_t match_id_base, uint32_t* unit_actions, uint8_t* tile_actions, uint64_t* counters,
                       void* cuda_stream);

/*
 * Episode statistics summed over the batch into out[16] (device int64):
 * [0] live matches, [1] finished, [2] wins A, [3] wins B, [4] ties, [5] sum turn,
 * [6] sum fuelGenerated, [7] sum units, [8] sum city tiles, [9] sum cities,
 * [10] matches with status != 0, [11] algorithmic bytes of one turn over live matches.
 * This is the vector NCCL all-reduces across GPUs.
 */
int lux_b200_stats(const LuxBatch* b, int64_t* out16, void* cuda_stream);

/*
 * Row offsets of the packed observation layout, computed on the device: offsets[m] = min(row_cap, sum over the
 * matches before m of (units + city tiles of the match's selected team)), offsets[n] = min(row_cap, total).
 * offsets: int64[n+1]; overflow (optional, int64*): the rows cut by row_cap are added to it.  Nothing is read back
 * to the host, so observe / encode of a turn can be captured in a CUDA graph with fixed-capacity row buffers.
 */
int lux_b200_observe_offsets(const LuxBatch* b, const uint8_t* team_sel, int64_t row_cap, int64_t* offsets,
                             int64_t* overflow, void* cuda_stream);

/*
 * End of an environment turn in one pass (LuxEnvironment.step's tail, luxai2021/env/lux_env.py:123-161, for every
 * env): reward as lux_b200_reward, done[m] = hdr.done, and for a finished match the reset from `pool` (entry chosen
 * like lux_b200_reset_done with epoch = *turn_dev), a cleared reward state, and the learning team of the new match
 * (MatchController.reset's coin flip, match_controller.py:118-122, as a counter-based hash of (match, *turn_dev)).
 * team_sel is updated in place.  reward: double[n], done: uint8[n], turn_dev: device uint32.
 */
int lux_b200_env_finish(const LuxBatch* b, const LuxBatch* pool, uint8_t* team_sel, int32_t* reward_state,
                        double* reward, uint8_t* done, const uint32_t* turn_dev, int32_t* reset_count, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LUX_B200_H */
