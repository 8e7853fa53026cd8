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
#include "lux_env_core.cuh"

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
        const uint32_t w = env_unit_word(units, n_units, S, slot, code);
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
