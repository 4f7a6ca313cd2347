// misc.cu -- Polynomial.scale / evaluate_domain on arbitrary points (code/univariate.py:145-169)
// and the small gathers behind Merkle.open / tree.leafs[i] (code/merkle.py:46-52, code/fri.py:150).
#include <vector>

#include "common.h"

namespace {

__device__ __forceinline__ xfe x_pow(xfe a, u64 e) {
    xfe acc = {{1, 0, 0}};
    while (e) {
        if (e & 1) acc = x_mul(acc, a);
        a = x_mul(a, a);
        e >>= 1;
    }
    return acc;
}

// code/univariate.py:168-169: c_i <- factor^i * c_i
template <int PLANES>
__global__ void __launch_bounds__(256)
    scale_kernel(const u64 *__restrict__ in, u64 in_stride, u64 *__restrict__ out, u64 out_stride, u64 n, u64 f0, u64 f1,
                 u64 f2) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (PLANES == 1) {
        out[i] = gl_mul(gl_pow(f0, i), in[i]);
    } else {
        const xfe f = x_pow(xfe{{f0, f1, f2}}, i);
        const xfe c = {{in[i], in[in_stride + i], in[2 * in_stride + i]}};
        const xfe r = x_mul(f, c);
        out[i] = r.c[0];
        out[out_stride + i] = r.c[1];
        out[2 * out_stride + i] = r.c[2];
    }
}

// code/univariate.py:145-151: value = sum_k c_k * x^k with a running power of the point
template <int CP, int PP>
__global__ void __launch_bounds__(128)
    eval_points_kernel(const u64 *__restrict__ coeffs, u64 cstride, u64 m, const u64 *__restrict__ pts, u64 pstride,
                       u64 npts, u64 *__restrict__ out, u64 ostride) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npts) return;
    if (CP == 1 && PP == 1) {
        const u64 x = pts[q];
        u64 xi = 1, val = 0;
        for (u64 k = 0; k < m; ++k) {
            val = gl_add(val, gl_mul(coeffs[k], xi));
            xi = gl_mul(xi, x);
        }
        out[q] = val;
    } else if (PP == 1) {  // extension coefficients, base-field point
        const u64 x = pts[q];
        u64 xi = 1;
        xfe val = {{0, 0, 0}};
        for (u64 k = 0; k < m; ++k) {
            const xfe c = {{coeffs[k], coeffs[cstride + k], coeffs[2 * cstride + k]}};
            val = x_add(val, x_mul_base(c, xi));
            xi = gl_mul(xi, x);
        }
        out[q] = val.c[0];
        out[ostride + q] = val.c[1];
        out[2 * ostride + q] = val.c[2];
    } else {
        const xfe x = {{pts[q], pts[pstride + q], pts[2 * pstride + q]}};
        xfe xi = {{1, 0, 0}}, val = {{0, 0, 0}};
        for (u64 k = 0; k < m; ++k) {
            xfe t;
            if (CP == 1) {
                t = x_mul_base(xi, coeffs[k]);
            } else {
                const xfe c = {{coeffs[k], coeffs[cstride + k], coeffs[2 * cstride + k]}};
                t = x_mul(c, xi);
            }
            val = x_add(val, t);
            xi = x_mul(xi, x);
        }
        out[q] = val.c[0];
        out[ostride + q] = val.c[1];
        out[2 * ostride + q] = val.c[2];
    }
}

__global__ void gather_kernel(const u64 *__restrict__ planes, u64 stride, u32 n_planes, const u64 *__restrict__ idx,
                              u32 n_idx, u64 *__restrict__ out) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * n_planes) return;
    const u32 q = t / n_planes, pl = t % n_planes;
    out[t] = planes[(u64)pl * stride + idx[q]];
}

// one thread per 16-byte quarter of one sibling digest
__global__ void open_kernel(const u8 *__restrict__ nodes, u64 npo2, u32 depth, const u64 *__restrict__ idx, u32 n_idx,
                            u8 *__restrict__ out) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * depth * 4) return;
    const u32 part = t & 3, j = (t >> 2) % depth, q = (t >> 2) / depth;
    const u64 k = ((npo2 | idx[q]) >> j) ^ 1;  // code/merkle.py:48-51
    reinterpret_cast<uint4 *>(out)[(size_t)(q * depth + j) * 4 + part] =
        reinterpret_cast<const uint4 *>(nodes)[k * 4 + part];
}

}  // namespace

extern "C" int b2s_scale(const uint64_t *d_in, uint64_t in_stride, uint64_t *d_out, uint64_t out_stride, uint64_t n,
                         uint32_t n_planes, const uint64_t factor[3], void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (n_planes == 1)
        scale_kernel<1><<<blocks, 256, 0, st>>>(d_in, in_stride, d_out, out_stride, n, factor[0], 0, 0);
    else if (n_planes == 3)
        scale_kernel<3><<<blocks, 256, 0, st>>>(d_in, in_stride, d_out, out_stride, n, factor[0], factor[1], factor[2]);
    else {
        b2s_set_error("scale: n_planes must be 1 or 3");
        return B2S_ERR_ARG;
    }
    B2S_LAUNCHED();
    return 0;
}

extern "C" int b2s_eval_points(const uint64_t *d_coeffs, uint64_t coeff_stride, uint32_t coeff_planes,
                               uint64_t n_coeffs, const uint64_t *d_points, uint64_t point_stride,
                               uint32_t point_planes, uint64_t n_points, uint64_t *d_out, uint64_t out_stride,
                               void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_points == 0) return 0;
    const unsigned blocks = (unsigned)((n_points + 127) / 128);
#define EV(CP, PP)                                                                                              \
    eval_points_kernel<CP, PP><<<blocks, 128, 0, st>>>(d_coeffs, coeff_stride, n_coeffs, d_points, point_stride, \
                                                       n_points, d_out, out_stride)
    if (coeff_planes == 1 && point_planes == 1)
        EV(1, 1);
    else if (coeff_planes == 3 && point_planes == 1)
        EV(3, 1);
    else if (coeff_planes == 1 && point_planes == 3)
        EV(1, 3);
    else if (coeff_planes == 3 && point_planes == 3)
        EV(3, 3);
    else {
        b2s_set_error("eval_points: planes must be 1 or 3");
        return B2S_ERR_ARG;
    }
#undef EV
    B2S_LAUNCHED();
    return 0;
}

extern "C" int b2s_gather(const uint64_t *d_planes, uint64_t plane_stride, uint32_t n_planes,
                          const uint64_t *h_indices, uint32_t n_indices, uint64_t *h_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_indices == 0) return 0;
    u64 *d_idx = nullptr, *d_out = nullptr;
    const u32 total = n_indices * n_planes;
    B2S_CUDA(cudaMallocAsync(&d_idx, sizeof(u64) * n_indices, st));
    B2S_CUDA(cudaMallocAsync(&d_out, sizeof(u64) * total, st));
    B2S_CUDA(cudaMemcpyAsync(d_idx, h_indices, sizeof(u64) * n_indices, cudaMemcpyHostToDevice, st));
    gather_kernel<<<(total + 127) / 128, 128, 0, st>>>(d_planes, plane_stride, n_planes, d_idx, n_indices, d_out);
    B2S_LAUNCHED();
    B2S_CUDA(cudaMemcpyAsync(h_out, d_out, sizeof(u64) * total, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaFreeAsync(d_idx, st));
    B2S_CUDA(cudaFreeAsync(d_out, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int b2s_merkle_open(const uint8_t *d_nodes, uint64_t npo2, const uint64_t *h_indices, uint32_t n_indices,
                               uint8_t *h_paths, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const u32 depth = ilog2_u64(npo2);
    if (n_indices == 0 || depth == 0) return 0;
    for (u32 q = 0; q < n_indices; ++q)
        if (h_indices[q] >= npo2) {
            b2s_set_error("merkle_open: index %llu out of range", (unsigned long long)h_indices[q]);
            return B2S_ERR_ARG;
        }
    u64 *d_idx = nullptr;
    u8 *d_out = nullptr;
    const size_t bytes = (size_t)n_indices * depth * 64;
    B2S_CUDA(cudaMallocAsync(&d_idx, sizeof(u64) * n_indices, st));
    B2S_CUDA(cudaMallocAsync(&d_out, bytes, st));
    B2S_CUDA(cudaMemcpyAsync(d_idx, h_indices, sizeof(u64) * n_indices, cudaMemcpyHostToDevice, st));
    const u32 total = n_indices * depth * 4;
    open_kernel<<<(total + 127) / 128, 128, 0, st>>>(d_nodes, npo2, depth, d_idx, n_indices, d_out);
    B2S_LAUNCHED();
    B2S_CUDA(cudaMemcpyAsync(h_paths, d_out, bytes, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaFreeAsync(d_idx, st));
    B2S_CUDA(cudaFreeAsync(d_out, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    return 0;
}
