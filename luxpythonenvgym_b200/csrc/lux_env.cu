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

// ---------------------------------------------------------------- packed observation row offsets
// offsets[m] = rows before match m when every match gets room for the units + city tiles of its selected team
// (an upper bound of its actionable entities), clamped to row_cap; offsets[n] = total.  Two passes: block-local
// exclusive scans (1024 matches per block) with the block totals left in scratch, then every block adds the sum
// of the totals before it.
#define LUX_SCAN_BLOCK 1024
__global__ void __launch_bounds__(LUX_SCAN_BLOCK) lux_offsets_local_kernel(LuxBatch b, const uint8_t* __restrict__ team_sel,
                                                                          int64_t* __restrict__ offsets, long long* __restrict__ totals) {
    LUX_PDL_TRIGGER();
    LUX_PDL_WAIT();
    __shared__ int warp_sum[32];
    const int m = blockIdx.x * LUX_SCAN_BLOCK + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int v = 0;
    if (m < b.n_matches) {
        const LuxHeader* h = b.hdr + m;
        const int team = team_sel[m] & 1;
        v = h->done ? 0 : (int)h->n_team_units[team] + (int)h->n_team_tiles[team];
    }
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sum[lane], xs = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, xs, d); if (lane >= d) xs += y; }
        warp_sum[lane] = xs - w;                      // exclusive
        if (lane == 31) totals[blockIdx.x] = xs;
    }
    __syncthreads();
    if (m < b.n_matches) offsets[m] = (int64_t)(warp_sum[warp] + x - v);
}

__global__ void __launch_bounds__(LUX_SCAN_BLOCK) lux_offsets_finish_kernel(int n, int64_t row_cap, int64_t* __restrict__ offsets,
                                                                           const long long* __restrict__ totals, int n_blocks,
                                                                           long long* __restrict__ overflow) {
    LUX_PDL_TRIGGER();
    LUX_PDL_WAIT();
    __shared__ long long partial[32];
    __shared__ long long base_s, all_s;
    long long mine = 0, every = 0;
    for (int k = threadIdx.x; k < n_blocks; k += LUX_SCAN_BLOCK) {
        const long long t = totals[k];
        every += t;
        if (k < (int)blockIdx.x) mine += t;
    }
    for (int pass = 0; pass < 2; pass++) {
        long long x = pass ? every : mine;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long s = 0;
            for (int w = 0; w < LUX_SCAN_BLOCK / 32; w++) s += partial[w];
            if (pass) all_s = s; else base_s = s;
        }
        __syncthreads();
    }
    const int m = blockIdx.x * LUX_SCAN_BLOCK + threadIdx.x;
    if (m < n) {
        const long long o = offsets[m] + base_s;
        offsets[m] = o < row_cap ? o : row_cap;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        offsets[n] = all_s < row_cap ? all_s : row_cap;
        if (overflow && all_s > row_cap) *overflow += all_s - row_cap;
    }
}

// ---------------------------------------------------------------- end of an environment turn
// One warp per env: AgentPolicy's per-turn reward (as lux_reward_kernel), the done flag, and for a finished match
// the reset from the pool, a cleared reward state and the learning team of the new match -- the coin flip of
// MatchController.reset (match_controller.py:118-122) as bit 17 of a counter-based hash of (env, turn).
__global__ void lux_env_finish_kernel(LuxBatch b, LuxBatch pool, uint8_t* __restrict__ team_sel, int32_t* __restrict__ last,
                                      double* __restrict__ reward, uint8_t* __restrict__ done_out,
                                      const uint32_t* __restrict__ turn_dev, int* __restrict__ reset_count) {
    LUX_PDL_TRIGGER();
    LUX_PDL_WAIT();
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= b.n_matches) return;
    const LuxHeader* h = b.hdr + m;
    const int done = h->done != 0;
    const uint32_t turn = *turn_dev;
    if (lane == 0) {
        const int team = team_sel[m] & 1;
        const int units_now = h->n_team_units[team], tiles_now = h->n_team_tiles[team], fuel = h->fuel_generated[team];
        int32_t* L = last + (size_t)m * 4;
        double r = 0.0;                               // same summation order as the Python dict loop (:556-558)
        r += (double)(units_now - L[0]) * 0.05;
        r += (double)(tiles_now - L[1]) * 0.1;
        r += (double)(fuel - L[2]) / 20000.0;
        r += done ? (double)tiles_now : 0.0;
        reward[m] = r;
        done_out[m] = (uint8_t)done;
        if (done) {
            L[0] = 0; L[1] = 0; L[2] = 0; L[3] = 0;
            uint32_t hsh = (uint32_t)m * 0x9E3779B1u + turn * 0x85EBCA6Bu;
            hsh *= 0x2C1B3C6Du;
            team_sel[m] = (uint8_t)((hsh >> 17) & 1u);
        } else {
            L[0] = units_now; L[1] = tiles_now; L[2] = fuel;
        }
    }
    if (!done) return;
    __syncwarp();
    // the pool entry: the rule of lux_reset_done_kernel (lux_step.cu) with epoch = turn
    int src = 0;
    if (lane == 0) {
        uint32_t hsh = (uint32_t)m * 0x9E3779B1u + turn * 0x85EBCA6Bu;
        hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du; hsh ^= hsh >> 12;
        src = (int)(hsh % (uint32_t)pool.n_matches);
        if (reset_count) atomicAdd(reset_count, 1);
    }
    src = __shfl_sync(0xffffffffu, src, 0);
    const int ncell = b.size * b.size;
    const uint4* sp[5] = {(const uint4*)(pool.cells + (size_t)src * ncell), (const uint4*)(pool.units + (size_t)src * b.u_cap),
                          (const uint4*)(pool.cities + (size_t)src * b.c_cap), (const uint4*)(pool.tiles + (size_t)src * b.t_cap),
                          (const uint4*)(pool.res + (size_t)src * b.r_cap)};
    uint4* dp[5] = {(uint4*)(b.cells + (size_t)m * ncell), (uint4*)(b.units + (size_t)m * b.u_cap), (uint4*)(b.cities + (size_t)m * b.c_cap),
                    (uint4*)(b.tiles + (size_t)m * b.t_cap), (uint4*)(b.res + (size_t)m * b.r_cap)};
    const LuxHeader* ph = pool.hdr + src;
    const int pieces[5] = {ncell / 2, ph->n_units, ph->n_cities, (ph->n_tiles * 2 + 15) / 16, (ph->n_res * 4 + 15) / 16};
#pragma unroll 1
    for (int sct = 0; sct < 5; sct++)
        for (int i = lane; i < pieces[sct]; i += 32) dp[sct][i] = sp[sct][i];
    if (lane < LUX_HDR_BYTES / 16) ((uint4*)(b.hdr + m))[lane] = ((const uint4*)ph)[lane];
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

static long long* g_scan_totals[16];      // per device: block totals of the offset scan (2048 blocks = 2 M matches)

extern "C" int lux_b200_observe_offsets(const LuxBatch* b, const uint8_t* team_sel, int64_t row_cap, int64_t* offsets,
                                        int64_t* overflow, void* cuda_stream) {
    int rc = env_check(b);
    if (rc) return rc;
    if (!team_sel || !offsets || row_cap < 0) return LUX_E_ARG;
    const int n = b->n_matches, n_blocks = (n + LUX_SCAN_BLOCK - 1) / LUX_SCAN_BLOCK;
    if (n_blocks > 2048) return LUX_E_CAPACITY;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return LUX_E_CUDA;
    if (!g_scan_totals[dev] && cudaMalloc(&g_scan_totals[dev], 2048 * sizeof(long long)) != cudaSuccess) return LUX_E_CUDA;
    if (n == 0) return cudaMemsetAsync(offsets, 0, sizeof(int64_t), st) == cudaSuccess ? LUX_OK : LUX_E_CUDA;
    cudaError_t e = lux_launch_chain(lux_offsets_local_kernel, dim3(n_blocks), dim3(LUX_SCAN_BLOCK), 0, st, *b, team_sel, offsets,
                                     g_scan_totals[dev]);
    if (e == cudaSuccess)
        e = lux_launch_chain(lux_offsets_finish_kernel, dim3(n_blocks), dim3(LUX_SCAN_BLOCK), 0, st, n, row_cap, offsets,
                             (const long long*)g_scan_totals[dev], n_blocks, (long long*)overflow);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "lux_b200: offsets launch: %s\n", cudaGetErrorString(e)); return LUX_E_CUDA; }
    return LUX_OK;
}

extern "C" int lux_b200_env_finish(const LuxBatch* b, const LuxBatch* pool, uint8_t* team_sel, int32_t* reward_state,
                                   double* reward, uint8_t* done, const uint32_t* turn_dev, int32_t* reset_count,
                                   void* cuda_stream) {
    int rc = env_check(b);
    if (rc) return rc;
    rc = env_check(pool);
    if (rc) return rc;
    if (!team_sel || !reward_state || !reward || !done || !turn_dev || pool->n_matches < 1) return LUX_E_ARG;
    if (pool->size != b->size || pool->u_cap != b->u_cap || pool->c_cap != b->c_cap || pool->t_cap != b->t_cap ||
        pool->r_cap != b->r_cap || !b->cells || !b->cities || !b->tiles || !b->res || !pool->cells || !pool->cities ||
        !pool->tiles || !pool->res)
        return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    const int wpc = 4;
    cudaError_t e = lux_launch_chain(lux_env_finish_kernel, dim3((b->n_matches + wpc - 1) / wpc), dim3(32 * wpc), 0,
                                     (cudaStream_t)cuda_stream, *b, *pool, team_sel, reward_state, reward, done, turn_dev,
                                     (int*)reset_count);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "lux_b200: env finish launch: %s\n", cudaGetErrorString(e)); return LUX_E_CUDA; }
    return LUX_OK;
}
