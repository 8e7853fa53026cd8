// lux_launch.cuh -- programmatic dependent launch (griddepcontrol) helpers shared by the kernels.
//
// Every kernel of the per-turn chain (action scatter / stream, classify, the two step classes, summary,
// reset, observe, encode, reward) lets its successor on the stream be scheduled early (LUX_PDL_TRIGGER, first
// thing) and waits for its predecessor's completion and memory before it touches anything (LUX_PDL_WAIT).  The
// successor's launch latency and ramp-up then hide behind the predecessor's tail.  Both instructions are
// no-ops in a kernel that was launched the ordinary way, and a predecessor that never triggers (any foreign
// kernel) simply releases its dependents when it completes.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#define LUX_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define LUX_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

extern "C" int lux_b200_overlap_enabled(void);      // the "overlap" option (lux_step.cu)

// launch as a programmatic dependent of the preceding kernel on the stream (ordinary launch with overlap off)
template <typename... KArgs, typename... Args>
static cudaError_t lux_launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = lux_b200_overlap_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
