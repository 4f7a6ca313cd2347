// common.h -- internal declarations shared by the translation units of libb2s.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/b2s.h"
#include "gl64.cuh"

extern std::atomic<uint64_t> g_launches;
void b2s_set_error(const char *fmt, ...);

#define B2S_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            b2s_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
            return e_ == cudaErrorMemoryAllocation ? B2S_ERR_NOMEM : B2S_ERR_CUDA;                \
        }                                                                                         \
    } while (0)

// count + check a kernel launch
#define B2S_LAUNCHED()                                                                            \
    do {                                                                                          \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                       \
        B2S_CUDA(cudaGetLastError());                                                             \
    } while (0)

static inline uint32_t ilog2_u64(uint64_t v) {
    uint32_t r = 0;
    while (v >>= 1) ++r;
    return r;
}

// ---- internal device-side entry points used across files -------------------------------
// ntt.cu
int ntt_run(const u64 *d_in, u64 in_stride, u32 n_in, u64 *d_out, u64 out_stride, u32 log_n, u32 n_planes, u64 omega,
            u64 offset, int inverse, cudaStream_t st);
void ntt_cache_clear();
// merkle.cu
int merkle_field_run(const u64 *d_planes, u64 stride, u64 n, const b2s_leaf_templates *tpl, u8 *d_nodes,
                     cudaStream_t st);
int merkle_upper_run(u8 *d_nodes, u64 npo2, cudaStream_t st);
int merkle_upload_templates(const b2s_leaf_templates *tpl, cudaStream_t st);

// ---- programmatic dependent launch (chains of small dependent kernels: tree levels, FRI rounds) ---------------
// A kernel launched with launch_pdl() may start while its predecessor in the stream is still running; it must
// execute pdl_wait() before it touches anything an earlier kernel wrote.  What precedes the wait (index arithmetic,
// staging constant tables) overlaps the predecessor's tail, and the launch latency disappears behind it.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const char *off = getenv("B2S_PDL");  // B2S_PDL=0: plain stream order (debugging)
    cfg.numAttrs = (off && off[0] == '0') ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
