// lux_obs.cu -- lux_b200_observe: per-entity observation tensors (see lux_obs_core.cuh).
// One warp per match; header, resource list and the leading unit / tile records are staged in shared
// memory by one cp.async.bulk (TMA) batch (a second batch only for populous matches); the cell grid and the
// cities are read in place; rows are zero-filled with coalesced 16-byte stores and the non-zero features
// written by the owning lanes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "lux_obs_core.cuh"
#include "lux_tma.cuh"
#include "lux_launch.cuh"

#define LUX_OBS_PRE_UNITS 32
#define LUX_OBS_PRE_TILES 32

struct LuxObsLayoutDev { LuxObsLayout L; int mbar; int total; };

template <int kS>
__global__ void lux_observe_kernel(LuxBatch b, const uint8_t* __restrict__ team_sel, float* __restrict__ obs,
                                   int16_t* __restrict__ ent_slot, int32_t* __restrict__ ent_count, int e_cap,
                                   const int64_t* __restrict__ row_offsets, LuxObsLayoutDev D) {
    extern __shared__ __align__(128) unsigned char lux_smem[];
    LUX_PDL_TRIGGER();
    LUX_PDL_WAIT();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= b.n_matches) return;
    unsigned char* smem = lux_smem + (size_t)warp * D.total;
    const int ncell = b.size * b.size;
    LuxObsMatch M;
    lux_obs_bind(M, smem, D.L, b.size, b.cells + (size_t)m * ncell, b.cities + (size_t)m * b.c_cap);
    uint64_t* bar = (uint64_t*)(smem + D.mbar);
    const LuxUnit* g_units = b.units + (size_t)m * b.u_cap;
    const uint16_t* g_tiles = b.tiles + (size_t)m * b.t_cap;
    const uint32_t pu = min(b.u_cap, LUX_OBS_PRE_UNITS), pt = min(b.t_cap, LUX_OBS_PRE_TILES);
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, LUX_HDR_BYTES + b.r_cap * 4 + pu * 16 + pt * 2);
        tma_g2s(M.hdr, b.hdr + m, LUX_HDR_BYTES, bar);
        tma_g2s(M.res, b.res + (size_t)m * b.r_cap, b.r_cap * 4, bar);
        tma_g2s(M.units, g_units, pu * 16, bar);
        tma_g2s(M.tiles, g_tiles, pt * 2, bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    // padded layout: rows of match m start at m * e_cap; packed layout: at row_offsets[m], with room
    // row_offsets[m + 1] - row_offsets[m] (the caller's upper bound, e.g. units + tiles of the team)
    const size_t row0 = row_offsets ? (size_t)row_offsets[m] : (size_t)m * e_cap;
    const int room = row_offsets ? (int)(row_offsets[m + 1] - row_offsets[m]) : e_cap;
    if (M.hdr->done) {
        if (lane == 0) ent_count[m] = 0;
        if (row_offsets) for (int i = lane; i < room; i += 32) ent_slot[row0 + i] = -1;
        return;
    }
    const uint32_t nu = M.hdr->n_units, nt = M.hdr->n_tiles;
    if (nu > pu || nt > pt) {
        const uint32_t ru = nu > pu ? nu - pu : 0, rt2 = nt > pt ? ((nt - pt) * 2 + 15) & ~15u : 0;
        if (lane == 0) {
            mbar_expect_tx(bar, ru * 16 + rt2);
            if (ru) tma_g2s(M.units + pu, g_units + pu, ru * 16, bar);
            if (rt2) tma_g2s(M.tiles + pt, g_tiles + pt, rt2, bar);
        }
        __syncwarp();
        mbar_wait(bar, 1);
    }
    int n = lux_observe<kS>(M, team_sel[m] & 1, obs + row0 * LUX_OBS_DIM, ent_slot + row0, room);
    if (n > room) n = room;
    if (lane == 0) ent_count[m] = n;
    if (row_offsets) for (int i = n + lane; i < room; i += 32) ent_slot[row0 + i] = -1;      // gap rows
}

extern "C" int lux_b200_device_count(void);

extern "C" int lux_b200_observe(const LuxBatch* b, const uint8_t* team_sel, float* obs, int16_t* ent_slot,
                                int32_t* ent_count, int e_cap, const int64_t* row_offsets, void* cuda_stream) {
    if (!b || !team_sel || !obs || !ent_slot || !ent_count || (e_cap < 1 && !row_offsets)) return LUX_E_ARG;
    if (b->n_matches < 0 || b->size < 4 || b->size > 32 || (b->size & 1)) return LUX_E_ARG;
    if (b->u_cap < 4 || b->u_cap > 252 || (b->u_cap & 3) || b->c_cap < 1 || b->c_cap > 255 || b->t_cap < 16 ||
        b->t_cap > 1024 || (b->t_cap & 15) || b->r_cap < 8 || b->r_cap > 1024 || (b->r_cap & 7))
        return LUX_E_CAPACITY;
    if (!b->hdr || !b->cells || !b->units || !b->cities || !b->tiles || !b->res) return LUX_E_ARG;
    if (b->n_matches == 0) return LUX_OK;
    if (lux_b200_device_count() <= 0) return LUX_E_NO_DEVICE;
    LuxObsLayoutDev D;
    D.L = lux_obs_layout(b->size, b->u_cap, b->t_cap, b->r_cap);
    D.mbar = D.L.total;
    D.total = D.L.total + 16;
    int wpc = D.total >= 8 * 1024 ? 1 : 2;
    size_t smem = (size_t)D.total * wpc;
    if (smem > 227 * 1024) return LUX_E_CAPACITY;
    cudaError_t e = cudaSuccess;
    auto launch = [&](auto kernel) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return;
        e = lux_launch_chain(kernel, dim3((b->n_matches + wpc - 1) / wpc), dim3(32 * wpc), smem, (cudaStream_t)cuda_stream,
                             *b, team_sel, obs, ent_slot, ent_count, e_cap, row_offsets, D);
    };
    switch (b->size) {           // the standard map sides get their own instantiation
        case 12: launch(lux_observe_kernel<12>); break;
        case 16: launch(lux_observe_kernel<16>); break;
        case 24: launch(lux_observe_kernel<24>); break;
        case 32: launch(lux_observe_kernel<32>); break;
        default: launch(lux_observe_kernel<0>); break;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "lux_b200: lux_observe_kernel launch: %s\n", cudaGetErrorString(e)); return LUX_E_CUDA; }
    return LUX_OK;
}
