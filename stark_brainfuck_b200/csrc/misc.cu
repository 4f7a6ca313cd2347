// misc.cu -- Polynomial.scale / evaluate_domain on arbitrary points (code/univariate.py:145-169)
// and the small gathers behind Merkle.open / tree.leafs[i] (code/merkle.py:46-52, code/fri.py:150).
#include <vector>

#include "common.h"
#include "glmont.cuh"

namespace {

GL_HD xfe x_pow(xfe a, u64 e) {
    xfe acc = {{1, 0, 0}};
    while (e) {
        if (e & 1) acc = x_mul(acc, a);
        a = x_mul(a, a);
        e >>= 1;
    }
    return acc;
}

// code/univariate.py:168-169: c_i <- factor^i * c_i.  A thread owns K coefficients i = t + k*T: one power
// factor^t, then one multiplication by factor^T per step (the reference's `factor ^ i` per coefficient is
// ~2 log2(i) multiplications each).
#define SCALE_K 8
template <int PLANES>
__global__ void __launch_bounds__(256)
    scale_kernel(const u64 *__restrict__ in, u64 in_stride, u64 *__restrict__ out, u64 out_stride, u64 n, u64 T, xfe f,
                 xfe step) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    if (PLANES == 1) {
        u64 p = gl_pow(f.c[0], t);
#pragma unroll
        for (int k = 0; k < SCALE_K; ++k) {
            const u64 i = t + k * T;
            if (i < n) out[i] = gl_mul(p, in[i]);
            p = gl_mul(p, step.c[0]);
        }
    } else {
        xfe p = x_pow(f, t);
#pragma unroll
        for (int k = 0; k < SCALE_K; ++k) {
            const u64 i = t + k * T;
            if (i < n) {
                const xfe c = {{in[i], in[in_stride + i], in[2 * in_stride + i]}};
                const xfe r = x_mul(p, c);
                out[i] = r.c[0];
                out[out_stride + i] = r.c[1];
                out[2 * out_stride + i] = r.c[2];
            }
            p = x_mul(p, step);
        }
    }
}

// code/univariate.py:145-151: value = sum_k c_k * x^k.  The reference walks the coefficients with a
// running power of the point; field arithmetic is exact, so any evaluation order gives the same
// canonical value.  Here the polynomial is cut into chunks: thread (point q, chunk ch) evaluates its
// chunk by Horner's rule (lanes of a warp share the coefficient loads), multiplies by x^(first index
// of the chunk) and a second kernel adds the partial values of a point.
template <int PP>
struct EvalAcc;
template <>
struct EvalAcc<1> {  // base-field point
    u64 x, xm;        // the point and its Montgomery form: the Horner steps are mont_mul + lazy add (glmont.cuh)
    __device__ __forceinline__ void load(const u64 *pts, u64, u64 q) {
        x = pts[q];
        xm = gl_to_mont(x);
    }
    __device__ __forceinline__ u64 mul(u64 v) const { return mont_mul(v, xm); }
    __device__ __forceinline__ xfe mul(const xfe &v) const {
        return {{mont_mul(v.c[0], xm), mont_mul(v.c[1], xm), mont_mul(v.c[2], xm)}};
    }
    __device__ __forceinline__ u64 times_power(u64 v, u64 e) const { return gl_mul(v, gl_pow(x, e)); }
    __device__ __forceinline__ xfe times_power(const xfe &v, u64 e) const { return x_mul_base(v, gl_pow(x, e)); }
};
template <>
struct EvalAcc<3> {  // extension-field point
    xfe x, xm;
    __device__ __forceinline__ void load(const u64 *pts, u64 pstride, u64 q) {
        x = xfe{{pts[q], pts[pstride + q], pts[2 * pstride + q]}};
        xm = xfe{{gl_to_mont(x.c[0]), gl_to_mont(x.c[1]), gl_to_mont(x.c[2])}};
    }
    __device__ __forceinline__ xfe mul(const xfe &v) const { return x_mul_mont(v, xm); }
    __device__ __forceinline__ xfe times_power(const xfe &v, u64 e) const { return x_mul(v, x_pow(x, e)); }
};

template <int CP, int PP>
__global__ void __launch_bounds__(128)
    eval_chunk_kernel(const u64 *__restrict__ coeffs, u64 cstride, u64 m, const u64 *__restrict__ pts, u64 pstride,
                      u64 npts, u64 chunk_len, u64 *__restrict__ partial) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npts) return;
    const u64 k0 = (u64)blockIdx.y * chunk_len;
    const u64 k1 = k0 + chunk_len < m ? k0 + chunk_len : m;
    EvalAcc<PP> P;
    P.load(pts, pstride, q);
    u64 *dst = partial + (u64)blockIdx.y * 3 * npts + q;
    if constexpr (CP == 1 && PP == 1) {
        u64 val = 0;  // lazy between the steps
        for (u64 k = k1; k-- > k0;) val = ladd(P.mul(val), coeffs[k]);
        dst[0] = P.times_power(lcanon(val), k0);
        dst[npts] = 0;
        dst[2 * npts] = 0;
    } else {
        xfe val = {{0, 0, 0}};
        for (u64 k = k1; k-- > k0;) {
            val = P.mul(val);
            val.c[0] = ladd(val.c[0], coeffs[k]);
            if (CP == 3) {
                val.c[1] = ladd(val.c[1], coeffs[cstride + k]);
                val.c[2] = ladd(val.c[2], coeffs[2 * cstride + k]);
            }
        }
        val = P.times_power(xfe{{lcanon(val.c[0]), lcanon(val.c[1]), lcanon(val.c[2])}}, k0);
        dst[0] = val.c[0];
        dst[npts] = val.c[1];
        dst[2 * npts] = val.c[2];
    }
}

__global__ void __launch_bounds__(128)
    eval_sum_kernel(const u64 *__restrict__ partial, u64 npts, u32 nchunks, u32 out_planes, u64 *__restrict__ out,
                    u64 ostride) {
    const u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npts) return;
    for (u32 pl = 0; pl < out_planes; ++pl) {
        u64 acc = 0;
        for (u32 ch = 0; ch < nchunks; ++ch) acc = gl_add(acc, partial[((u64)ch * 3 + pl) * npts + q]);
        out[(u64)pl * ostride + q] = acc;
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
    const u64 T = (n + SCALE_K - 1) / SCALE_K;
    const unsigned blocks = (unsigned)((T + 255) / 256);
    const xfe f = {{factor[0] % GL_P, n_planes == 3 ? factor[1] % GL_P : 0, n_planes == 3 ? factor[2] % GL_P : 0}};
    const xfe step = x_pow(f, T);
    if (n_planes == 1)
        scale_kernel<1><<<blocks, 256, 0, st>>>(d_in, in_stride, d_out, out_stride, n, T, f, step);
    else if (n_planes == 3)
        scale_kernel<3><<<blocks, 256, 0, st>>>(d_in, in_stride, d_out, out_stride, n, T, f, step);
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
    if ((coeff_planes != 1 && coeff_planes != 3) || (point_planes != 1 && point_planes != 3)) {
        b2s_set_error("eval_points: planes must be 1 or 3");
        return B2S_ERR_ARG;
    }
    const u32 out_planes = coeff_planes > point_planes ? coeff_planes : point_planes;
    if (n_coeffs == 0) {  // the zero polynomial
        B2S_CUDA(cudaMemset2DAsync(d_out, sizeof(u64) * out_stride, 0, sizeof(u64) * n_points, out_planes, st));
        return 0;
    }
    const unsigned blocks = (unsigned)((n_points + 127) / 128);
    // enough (point, chunk) threads to fill the GPU, at least 64 coefficients per chunk
    u64 nchunks = ((u64)1 << 17) / n_points;
    const u64 max_chunks = (n_coeffs + 63) / 64;
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 65535) nchunks = 65535;
    const u64 chunk_len = (n_coeffs + nchunks - 1) / nchunks;
    nchunks = (n_coeffs + chunk_len - 1) / chunk_len;
    u64 *partial = nullptr;
    B2S_CUDA(cudaMallocAsync(&partial, sizeof(u64) * 3 * nchunks * n_points, st));
    const dim3 grid(blocks, (unsigned)nchunks);
#define EV(CP, PP)                                                                                                      \
    eval_chunk_kernel<CP, PP><<<grid, 128, 0, st>>>(d_coeffs, coeff_stride, n_coeffs, d_points, point_stride, n_points, \
                                                    chunk_len, partial)
    if (coeff_planes == 1 && point_planes == 1)
        EV(1, 1);
    else if (coeff_planes == 3 && point_planes == 1)
        EV(3, 1);
    else if (coeff_planes == 1 && point_planes == 3)
        EV(1, 3);
    else
        EV(3, 3);
#undef EV
    B2S_LAUNCHED();
    eval_sum_kernel<<<blocks, 128, 0, st>>>(partial, n_points, (u32)nchunks, out_planes, d_out, out_stride);
    B2S_LAUNCHED();
    cudaFreeAsync(partial, st);
    return 0;
}

// ---- everything the query phase of one FRI proof opens, in one launch and one synchronisation ------------
namespace {
struct OpenItem {
    const u64 *planes;  // or null
    const u8 *nodes;    // or null
    u64 stride, npo2, index, path_off;
    u32 depth, pad;
};

// one CTA per opened index: thread t < n_planes copies one coefficient, thread t < 4 * depth one 16-byte quarter
// of the sibling digest at level t / 4 (code/merkle.py:46-52)
__global__ void __launch_bounds__(128)
    open_multi_kernel(const OpenItem *__restrict__ items, u32 n_planes, u64 *__restrict__ values, u8 *__restrict__ paths) {
    const OpenItem it = items[blockIdx.x];
    const u32 t = threadIdx.x;
    if (it.planes && t < n_planes) values[(u64)blockIdx.x * n_planes + t] = it.planes[(u64)t * it.stride + it.index];
    if (it.nodes && t < 4 * it.depth) {
        const u32 level = t >> 2, quarter = t & 3;
        const u64 k = ((it.npo2 + it.index) >> level) ^ 1;
        const uint4 v = *reinterpret_cast<const uint4 *>(it.nodes + k * 64 + quarter * 16);
        *reinterpret_cast<uint4 *>(paths + it.path_off + (u64)level * 64 + quarter * 16) = v;
    }
}
}  // namespace

extern "C" int b2s_open_multi(const uint64_t *const *h_planes, const uint64_t *h_plane_strides, uint32_t n_planes,
                              const uint8_t *const *h_nodes, const uint64_t *h_npo2, const uint32_t *h_counts,
                              const uint64_t *h_indices, uint32_t n_sets, uint64_t *h_values, uint8_t *h_paths,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_planes == 0 || n_planes > 128) {
        b2s_set_error("open_multi: n_planes %u out of range", n_planes);
        return B2S_ERR_ARG;
    }
    std::vector<OpenItem> items;
    u64 path_bytes = 0, pos = 0;
    for (u32 s = 0; s < n_sets; ++s) {
        const u32 depth = h_nodes[s] ? ilog2_u64(h_npo2[s]) : 0;
        if (h_nodes[s] && (h_npo2[s] == 0 || (h_npo2[s] & (h_npo2[s] - 1)) || depth > 31)) {
            b2s_set_error("open_multi: tree %u has %llu leaf slots", s, (unsigned long long)h_npo2[s]);
            return B2S_ERR_ARG;
        }
        for (u32 q = 0; q < h_counts[s]; ++q, ++pos) {
            if (h_nodes[s] && h_indices[pos] >= h_npo2[s]) {
                b2s_set_error("open_multi: index %llu out of range", (unsigned long long)h_indices[pos]);
                return B2S_ERR_ARG;
            }
            OpenItem it{};
            it.planes = h_planes[s];
            it.nodes = h_nodes[s];
            it.stride = h_plane_strides[s];
            it.npo2 = h_npo2[s];
            it.index = h_indices[pos];
            it.path_off = path_bytes;
            it.depth = depth;
            items.push_back(it);
            path_bytes += (u64)depth * 64;
        }
    }
    if (items.empty()) return 0;
    const size_t n = items.size(), vbytes = sizeof(u64) * n * n_planes;
    u8 *d_buf = nullptr;  // [items | values | paths]
    const size_t ibytes = (sizeof(OpenItem) * n + 15) / 16 * 16, vpad = (vbytes + 15) / 16 * 16;
    B2S_CUDA(cudaMallocAsync(&d_buf, ibytes + vpad + path_bytes + 16, st));
    B2S_CUDA(cudaMemcpyAsync(d_buf, items.data(), sizeof(OpenItem) * n, cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemsetAsync(d_buf + ibytes, 0, vpad, st));  // sets without planes read as zeros
    open_multi_kernel<<<(unsigned)n, 128, 0, st>>>(reinterpret_cast<const OpenItem *>(d_buf), n_planes,
                                                   reinterpret_cast<u64 *>(d_buf + ibytes), d_buf + ibytes + vpad);
    B2S_LAUNCHED();
    if (h_values) B2S_CUDA(cudaMemcpyAsync(h_values, d_buf + ibytes, vbytes, cudaMemcpyDeviceToHost, st));
    if (h_paths && path_bytes)
        B2S_CUDA(cudaMemcpyAsync(h_paths, d_buf + ibytes + vpad, path_bytes, cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaFreeAsync(d_buf, st));
    B2S_CUDA(cudaStreamSynchronize(st));
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
