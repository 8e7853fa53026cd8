// lux_env.cu -- the LuxEnvironment-facing pieces around the turn engine:
//   lux_b200_encode_actions : AgentPolicy action codes -> packed action tensors
//                             (examples/agent_policy.py:117-132, :426-470, smart_transfer_to_nearby :24-100)
//   lux_b200_reward         : AgentPolicy.get_reward (examples/agent_policy.py:494-560), once per turn
// Both read the match in place from global memory (a few hundred bytes per match are touched).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lux_b200.h"
#include "lux_launch.cuh"

#define ENV_CAP_WORKER 100
#define ENV_CAP_CART 2000

__device__ __forceinline__ int env_space(const LuxUnit& u) {
    return ((u.team_type & 2) ? ENV_CAP_CART : ENV_CAP_WORKER) - ((int)u.wood + (int)u.coal + (int)u.uranium);
}

// smart_transfer_to_nearby: largest cargo type (dict order wood, uranium, coal; strict >), best unit of
// the wanted type on the N, E, S, W cells (units of one cell in slot order)
__device__ uint32_t env_smart_transfer(const LuxUnit* units, int n_units, int S, int i, int want_cart) {
    const LuxUnit U = units[i];
    const int team = U.team_type & 1;
    const int amounts[3] = {U.wood, U.uranium, U.coal};
    const int codes[3] = {0, 2, 1};
    int rtype = -1, ramount = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (amounts[k] > ramount) { rtype = codes[k]; ramount = amounts[k]; }
    int target = -1, st = 0;
    for (int k = 0; k < 4; k++) {
        int cx = U.x + (k == 1 ? 1 : k == 3 ? -1 : 0), cy = U.y + (k == 0 ? -1 : k == 2 ? 1 : 0);
        if (cx < 0 || cy < 0 || cx >= S || cy >= S) continue;
        for (int j = 0; j < n_units; j++) {
            const LuxUnit V = units[j];
            if (V.x != cx || V.y != cy || ((V.team_type >> 1) & 1) != want_cart || (V.team_type & 1) != team) continue;
            int sv = env_space(V);
            if (target < 0) { target = j; st = sv; continue; }
            if (sv >= ramount && st >= ramount) { if (sv < st) { target = j; st = sv; } }
            else if (st >= ramount) { }
            else if (sv > st) { target = j; st = sv; }
        }
    }
    if (target < 0 || rtype < 0) return 0;
    if (st < ramount) ramount = st;
    if (ramount > 2047) ramount = 2047;
    return LUX_UA_TRANSFER | ((uint32_t)rtype << 3) | ((uint32_t)ramount << 6) | ((uint32_t)units[target].id << 17);
}

__global__ void lux_encode_actions_kernel(LuxBatch b, const int16_t* __restrict__ ent_slot,
                                          const int32_t* __restrict__ ent_count, const int32_t* __restrict__ codes,
                                          int e_cap, const int64_t* __restrict__ row_offsets,
                                          uint32_t* __restrict__ unit_actions, uint8_t* __restrict__ tile_actions) {
    LUX_PDL_TRIGGER();
    LUX_PDL_WAIT();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= b.n_matches) return;
    uint32_t* ua = unit_actions + (size_t)m * b.u_cap;
    uint8_t* ta = tile_actions + (size_t)m * b.t_cap;
    for (int i = lane; i < b.u_cap; i += 32) ua[i] = 0;
    for (int i = lane; i < b.t_cap / 4; i += 32) ((uint32_t*)ta)[i] = 0;
    __syncwarp();
    const LuxHeader* h = b.hdr + m;
    if (h->done) return;
    const LuxUnit* units = b.units + (size_t)m * b.u_cap;
    const int n_units = h->n_units, S = b.size;
    const size_t row0 = row_offsets ? (size_t)row_offsets[m] : (size_t)m * e_cap;
    const int room = row_offsets ? (int)(row_offsets[m + 1] - row_offsets[m]) : e_cap;
    int n = ent_count[m];
    if (n > room) n = room;
    for (int e = lane; e < n; e += 32) {
        int code = codes[row0 + e], slot = ent_slot[row0 + e];
        if (slot & 0x4000) {
            int p = slot & 0x3FFF;
            if (p < b.t_cap) ta[p] = (uint8_t)(1 + ((code % 3) + 3) % 3);
            continue;
        }
        if (slot < 0 || slot >= n_units) continue;
        int c = ((code % 9) + 9) % 9;
        uint32_t w;
        if (c == 0) w = LUX_UA_MOVE | (LUX_DIR_CENTER << 3);
        else if (c == 1) w = LUX_UA_MOVE | (LUX_DIR_NORTH << 3);
        else if (c == 2) w = LUX_UA_MOVE | (LUX_DIR_WEST << 3);
        else if (c == 3) w = LUX_UA_MOVE | (LUX_DIR_SOUTH << 3);
        else if (c == 4) w = LUX_UA_MOVE | (LUX_DIR_EAST << 3);
        else if (c == 5) w = env_smart_transfer(units, n_units, S, slot, 1);
        else if (c == 6) w = env_smart_transfer(units, n_units, S, slot, 0);
        else if (c == 7) w = LUX_UA_BCITY;
        else w = LUX_UA_PILLAGE;
        ua[slot] = w;
    }
}

__global__ void lux_reward_kernel(LuxBatch b, const uint8_t* __restrict__ team_sel, int32_t* __restrict__ last,
                                  double* __restrict__ reward) {
    LUX_PDL_TRIGGER();
    LUX_PDL_WAIT();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.n_matches) return;
    const LuxHeader* h = b.hdr + m;
    const int team = team_sel[m] & 1;
    const int units_now = h->n_team_units[team], tiles_now = h->n_team_tiles[team], fuel = h->fuel_generated[team];
    int32_t* L = last + (size_t)m * 4;
    double r = 0.0;                                   // same summation order as the Python dict loop (:556-558)
    r += (double)(units_now - L[0]) * 0.05;
    r += (double)(tiles_now - L[1]) * 0.1;
    r += (double)(fuel - L[2]) / 20000.0;
    r += h->done ? (double)tiles_now : 0.0;
    L[0] = units_now; L[1] = tiles_now; L[2] = fuel;
    reward[m] = r;
}

extern "C" int lux_b200_device_count(void);

static int env_check(const LuxBatch* b) {
    if (!b || b->n_matches < 0 || b->size < 4 || b->size > 32) return LUX_E_ARG;
    if (b->u_cap < 4 || b->t_cap < 16 || (b->t_cap & 15)) return LUX_E_CAPACITY;
    if (!b->hdr || !b->units) return LUX_E_ARG;
    return LUX_OK;
}

extern "C" int lux_b200_encode_actions(const LuxBatch* b, const int16_t* ent_slot, const int32_t* ent_count,
                                       const int32_t* codes, int e_cap, const int64_t* row_offsets,
                                       uint32_t* unit_actions, uint8_t* tile_actions, void* cuda_stream) {
    int rc = env_check(b);
    if (rc) return rc;
    if (!ent_slot || !ent_count || !codes || !unit_actions || !tile_actions || (e_cap < 1 && !row_offsets)) return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    const int wpc = 4;
    cudaError_t e = lux_launch_chain(lux_encode_actions_kernel, dim3((b->n_matches + wpc - 1) / wpc), dim3(32 * wpc), 0,
                                     (cudaStream_t)cuda_stream, *b, ent_slot, ent_count, codes, e_cap, row_offsets, unit_actions,
                                     tile_actions);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "lux_b200: encode launch: %s\n", cudaGetErrorString(e)); return LUX_E_CUDA; }
    return LUX_OK;
}

extern "C" int lux_b200_reward(const LuxBatch* b, const uint8_t* team_sel, int32_t* reward_state, double* reward,
                               void* cuda_stream) {
    int rc = env_check(b);
    if (rc) return rc;
    if (!team_sel || !reward_state || !reward) return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    cudaError_t e = lux_launch_chain(lux_reward_kernel, dim3((b->n_matches + 127) / 128), dim3(128), 0, (cudaStream_t)cuda_stream,
                                     *b, team_sel, reward_state, reward);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "lux_b200: reward launch: %s\n", cudaGetErrorString(e)); return LUX_E_CUDA; }
    return LUX_OK;
}
