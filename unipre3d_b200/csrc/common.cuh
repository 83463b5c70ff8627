// common.cuh -- shared helpers of libunipre3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/up3d.h"

#define UP3D_TILE 16            // BLOCK_X == BLOCK_Y of the reference rasterizer
#define UP3D_TILE_PIX 256
#define UP3D_NUM_SMS 148

namespace up3d {

// thread-local last error (exported through up3d_last_error)
char *err_buf();
int set_error(const char *fmt, ...);

#define UP3D_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) return up3d::set_error(__VA_ARGS__); \
    } while (0)

#define UP3D_CUDA_OK(expr)                                                                             \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return up3d::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)

#define UP3D_LAUNCH_OK(name)                                                                   \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) return up3d::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- exactly-rounded single operations: never contracted by the compiler.  The tile-assignment
// arithmetic is written with these so that it is bit-identical to oracle/raster_oracle.c. ----
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may begin while its predecessor on
// the stream is still draining; it must execute pdl_wait() before its first access to memory the predecessor (or any
// earlier kernel of the stream) touches.  pdl_trigger() lets the NEXT kernel of the stream start its own prologue.
// Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // UP3D_PDL=0 in the environment or up3d_set_pdl(0) turns the launch attribute off

template <typename... KArgs, typename... CallArgs>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     CallArgs... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace up3d
