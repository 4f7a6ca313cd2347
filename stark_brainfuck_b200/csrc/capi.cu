// capi.cu -- library management, error reporting and the NTT entry points of include/b2s.h
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.h"

std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";
static int g_sm_count = 0;

void b2s_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int b2s_version(void) { return 1; }
extern "C" const char *b2s_last_error(void) { return g_err; }
extern "C" uint64_t b2s_launch_count(void) { return g_launches.load(); }
extern "C" int b2s_device_sm_count(void) { return g_sm_count; }

extern "C" uint64_t b2s_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a % GL_P, b % GL_P); }
extern "C" uint64_t b2s_gl_pow(uint64_t a, uint64_t e) { return gl_pow(a % GL_P, e); }
extern "C" uint64_t b2s_gl_inv(uint64_t a) { return gl_inv(a % GL_P); }

extern "C" int b2s_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        b2s_set_error("no CUDA device visible (%s); libb2s has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return B2S_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        b2s_set_error("device %d out of range (%d visible)", device, count);
        return B2S_ERR_ARG;
    }
    B2S_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2S_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        b2s_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return B2S_ERR_CUDA;
    }
    g_sm_count = prop.multiProcessorCount;
    // Keep freed scratch (the transform's work buffer, gather buffers) in the stream-ordered pool instead of
    // returning it to the driver at every synchronisation -- up to a bound: the caller's own allocator (torch's
    // caching allocator uses cudaMalloc) shares the device, and what this pool holds back it cannot use.
    // b2s_trim() returns everything.
    cudaMemPool_t pool;
    B2S_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = (uint64_t)2 << 30;
    if (const char *mb = getenv("B2S_POOL_THRESHOLD_MB")) thresh = (uint64_t)strtoull(mb, nullptr, 10) << 20;
    B2S_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    return 0;
}

extern "C" int b2s_trim(void) {
    int device = 0;
    B2S_CUDA(cudaGetDevice(&device));
    B2S_CUDA(cudaDeviceSynchronize());
    cudaMemPool_t pool;
    B2S_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    B2S_CUDA(cudaMemPoolTrimTo(pool, 0));
    ntt_cache_clear();
    return 0;
}

// The host-buffer path keeps one pair of device staging buffers and one non-blocking stream
// per thread (grown on demand, released by b2s_shutdown of that thread or at exit): a malloc/free
// pair per call would cost more than the transform.
namespace {
struct HostPath {
    cudaStream_t st = nullptr;
    u64 *d_in = nullptr, *d_out = nullptr;
    size_t cap_in = 0, cap_out = 0;
    int dev = -1;
};
thread_local HostPath g_hp;

int host_path_reserve(size_t n_in_elems, size_t n_out_elems) {
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    if (g_hp.dev != dev) {
        if (g_hp.st) cudaStreamDestroy(g_hp.st);
        if (g_hp.d_in) cudaFree(g_hp.d_in);
        if (g_hp.d_out) cudaFree(g_hp.d_out);
        g_hp = HostPath();
        g_hp.dev = dev;
        B2S_CUDA(cudaStreamCreateWithFlags(&g_hp.st, cudaStreamNonBlocking));
    }
    if (n_in_elems > g_hp.cap_in) {
        if (g_hp.d_in) cudaFree(g_hp.d_in);
        g_hp.d_in = nullptr;
        B2S_CUDA(cudaMalloc(&g_hp.d_in, sizeof(u64) * n_in_elems));
        g_hp.cap_in = n_in_elems;
    }
    if (n_out_elems > g_hp.cap_out) {
        if (g_hp.d_out) cudaFree(g_hp.d_out);
        g_hp.d_out = nullptr;
        B2S_CUDA(cudaMalloc(&g_hp.d_out, sizeof(u64) * n_out_elems));
        g_hp.cap_out = n_out_elems;
    }
    return 0;
}

int copy_planes(u64 *dst, u64 dst_stride, const u64 *src, u64 src_stride, u64 width, u32 n_planes, cudaMemcpyKind kind,
                cudaStream_t st) {
    if (n_planes == 1 || (dst_stride == width && src_stride == width)) {
        B2S_CUDA(cudaMemcpyAsync(dst, src, sizeof(u64) * width * n_planes, kind, st));
    } else {
        B2S_CUDA(cudaMemcpy2DAsync(dst, sizeof(u64) * dst_stride, src, sizeof(u64) * src_stride, sizeof(u64) * width,
                                   n_planes, kind, st));
    }
    return 0;
}
}  // namespace

extern "C" int b2s_shutdown(void) {
    cudaDeviceSynchronize();
    if (g_hp.st) cudaStreamDestroy(g_hp.st);
    if (g_hp.d_in) cudaFree(g_hp.d_in);
    if (g_hp.d_out) cudaFree(g_hp.d_out);
    g_hp = HostPath();
    ntt_cache_clear();
    return 0;
}

extern "C" int b2s_ntt(const uint64_t *d_in, uint64_t in_stride, uint32_t n_in, uint64_t *d_out, uint64_t out_stride,
                       uint32_t log_n, uint32_t n_planes, uint64_t omega, uint64_t offset, int inverse, void *stream) {
    return ntt_run(d_in, in_stride, n_in, d_out, out_stride, log_n, n_planes, omega, offset, inverse,
                   (cudaStream_t)stream);
}

extern "C" int b2s_ntt_host(const uint64_t *h_in, uint64_t in_stride, uint32_t n_in, uint64_t *h_out,
                            uint64_t out_stride, uint32_t log_n, uint32_t n_planes, uint64_t omega, uint64_t offset,
                            int inverse) {
    if (log_n > 30) {
        b2s_set_error("log_n %u too large", log_n);
        return B2S_ERR_ARG;
    }
    const u64 n = (u64)1 << log_n;
    int rc = host_path_reserve((size_t)(n_in ? n_in : 1) * n_planes, (size_t)n * n_planes);
    if (rc) return rc;
    cudaStream_t st = g_hp.st;
    if (n_in) {
        rc = copy_planes(g_hp.d_in, n_in, h_in, in_stride, n_in, n_planes, cudaMemcpyHostToDevice, st);
        if (rc) return rc;
    }
    rc = ntt_run(g_hp.d_in, n_in, n_in, g_hp.d_out, n, log_n, n_planes, omega, offset, inverse, st);
    if (rc == 0) rc = copy_planes(h_out, out_stride, g_hp.d_out, n, n, n_planes, cudaMemcpyDeviceToHost, st);
    B2S_CUDA(cudaStreamSynchronize(st));
    return rc;
}

extern "C" int b2s_ntt_timed(const uint64_t *d_in, uint64_t in_stride, uint32_t n_in, uint64_t *d_out,
                             uint64_t out_stride, uint32_t log_n, uint32_t n_planes, uint64_t omega, uint64_t offset,
                             int inverse, void *stream, uint32_t iters, float *ms) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    B2S_CUDA(cudaEventCreate(&e0));
    B2S_CUDA(cudaEventCreate(&e1));
    B2S_CUDA(cudaEventRecord(e0, st));
    int rc = 0;
    for (uint32_t i = 0; i < iters && rc == 0; ++i)
        rc = ntt_run(d_in, in_stride, n_in, d_out, out_stride, log_n, n_planes, omega, offset, inverse, st);
    B2S_CUDA(cudaEventRecord(e1, st));
    B2S_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    B2S_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms) *ms = iters ? t / iters : 0.f;
    return rc;
}
