// combine.cu -- the nonlinear combination codeword.
//
// Replaces the inline block code/brainfuck_stark.py:241-298: for every base, extension and
// quotient codeword c (151 of them for the Brainfuck AIR) the reference builds the list c and the
// degree-shifted list x_j^shift * c[j], multiplies each by a Fiat-Shamir weight and adds everything
// up element by element in Python.  As field arithmetic that is
//     out[j] = sum_c (wa_c + wb_c * x_j^shift_c) * c[j],      x_j = offset * omega^j,
// which one kernel evaluates with the codewords staying where the producing kernels left them
// (NTT outputs, quotient outputs): a column is any base-field plane or extension-field plane
// triple in device memory.  Columns are grouped by shift ("slot") so that x_j^shift is computed
// once per slot: a thread owns K points j = t + k*T and walks x^shift from one to the next with a
// single multiplication by omega^(shift*T).
//
// Roofline: a point reads 8 B per base column and 24 B per extension column once (3.3 KB for the
// 151 columns of the Brainfuck AIR) and writes 24 B, against ~12 field multiplications per
// extension column -- ~1700 64-bit modular multiplications per point, so the kernel is bound by
// the integer pipes, not HBM.
#include <algorithm>
#include <vector>

#include "common.h"
#include "glmont.cuh"

namespace {

struct CombSlot {
    u64 off;      // offset^shift
    u64 step;     // omega^(shift * T)
    u64 sq[32];   // (omega^shift)^(2^b)
    u32 begin, end;
    u32 shifted;  // 0: columns of this slot have no shifted term (wb == 0)
    u32 pad;
};

struct CombCol {
    const u64 *ptr;
    u64 stride;
    u64 wa[3];  // wa * 2^64   (Montgomery form: mont_mul(value, w) is then the plain product)
    u64 wb[3];  // wb * 2^128  (so that mont_mul(x^shift, wb) is (wb * x^shift) * 2^64)
    u32 planes, pad;
};

// (wa + wb * x) * 2^64, canonical
__device__ __forceinline__ xfe comb_weight(const xfe &wa, const xfe &wb, bool shifted, u64 x) {
    if (!shifted) return wa;
    xfe w;
#pragma unroll
    for (int j = 0; j < 3; ++j) w.c[j] = lcanon(ladd(mont_mul(x, wb.c[j]), wa.c[j]));
    return w;
}

template <int K>
__global__ void __launch_bounds__(256)
    comb_kernel(const CombSlot *__restrict__ slots, u32 n_slots, const CombCol *__restrict__ cols, u64 T,
                u64 *__restrict__ out, u64 out_stride) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    xfe acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = {{0, 0, 0}};
    for (u32 d = 0; d < n_slots; ++d) {
        const CombSlot *S = slots + d;
        const bool shifted = S->shifted != 0;
        u64 x[K];
        if (shifted) {
            x[0] = gl_mul(S->off, gl_pow_sq(S->sq, t));
            const u64 step = S->step;
#pragma unroll
            for (int k = 1; k < K; ++k) x[k] = gl_mul(x[k - 1], step);
        }
        const u32 end = S->end;
        for (u32 c = S->begin; c < end; ++c) {
            const CombCol *C = cols + c;
            const u64 *p = C->ptr + t;
            const u64 stride = C->stride;
            const xfe wa = {{C->wa[0], C->wa[1], C->wa[2]}};
            const xfe wb = {{C->wb[0], C->wb[1], C->wb[2]}};
            if (C->planes == 3) {
                xfe v[K];
#pragma unroll
                for (int k = 0; k < K; ++k) v[k] = {{p[k * T], p[stride + k * T], p[2 * stride + k * T]}};
#pragma unroll
                for (int k = 0; k < K; ++k) x_fma_mont(acc[k], v[k], comb_weight(wa, wb, shifted, x[k]));
            } else {
                u64 v[K];
#pragma unroll
                for (int k = 0; k < K; ++k) v[k] = p[k * T];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const xfe w = comb_weight(wa, wb, shifted, x[k]);
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc[k].c[j] = ladd(acc[k].c[j], mont_mul(v[k], w.c[j]));
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        u64 *o = out + t + k * T;
        o[0] = lcanon(acc[k].c[0]);
        o[out_stride] = lcanon(acc[k].c[1]);
        o[2 * out_stride] = lcanon(acc[k].c[2]);
    }
}

}  // namespace

extern "C" int b2s_combination(const uint64_t *const *h_cols, const uint64_t *h_strides, const uint32_t *h_planes,
                               const uint64_t *h_wa, const uint64_t *h_wb, const uint64_t *h_shifts, uint32_t n_cols,
                               uint64_t N, uint64_t offset, uint64_t omega, uint64_t *d_out, uint64_t out_stride,
                               void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || (N & (N - 1)) || out_stride < N) {
        b2s_set_error("combination: bad arguments (N %llu, out_stride %llu)", (unsigned long long)N,
                      (unsigned long long)out_stride);
        return B2S_ERR_ARG;
    }
    for (u32 c = 0; c < n_cols; ++c)
        if ((h_planes[c] != 1 && h_planes[c] != 3) || !h_cols[c] || (h_planes[c] == 3 && h_strides[c] < N)) {
            b2s_set_error("combination: column %u has %u planes, stride %llu", c, h_planes[c],
                          (unsigned long long)h_strides[c]);
            return B2S_ERR_ARG;
        }
    // points per thread: 2 (64 registers, three CTAs per SM) beats 4 (96 registers) at every size measured on a
    // B200 -- 2^16: 0.25 vs 0.32 ms, 2^18: 0.46 vs 0.60, 2^20: 1.19 vs 1.21 for the AIR's 76 columns -- and 1
    static const char *force_k = getenv("B2S_COMB_K");
    const int K = N < 1024 ? 1 : (force_k ? atoi(force_k) : 2);
    if (K != 1 && K != 2 && K != 4) {
        b2s_set_error("combination: B2S_COMB_K must be 1, 2 or 4");
        return B2S_ERR_ARG;
    }
    const u64 T = N / K;
    // group the columns by shift; columns without a shifted term form slot 0
    std::vector<u32> order(n_cols);
    for (u32 c = 0; c < n_cols; ++c) order[c] = c;
    auto has_b = [&](u32 c) { return (h_wb[3 * c] | h_wb[3 * c + 1] | h_wb[3 * c + 2]) != 0; };
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) {
        const bool sa = has_b(a), sb = has_b(b);
        if (sa != sb) return !sa;
        return sa && h_shifts[a] < h_shifts[b];
    });
    std::vector<CombSlot> slots;
    std::vector<CombCol> cols(n_cols);
    for (u32 i = 0; i < n_cols; ++i) {
        const u32 c = order[i];
        const bool sh = has_b(c);
        const bool fresh = slots.empty() || (slots.back().shifted != 0) != sh ||
                           (sh && h_shifts[order[i - 1]] != h_shifts[c]);
        if (fresh) {
            CombSlot S{};
            S.begin = S.end = i;
            S.shifted = sh ? 1 : 0;
            if (sh) {
                const u64 s = h_shifts[c];
                S.off = gl_pow(offset % GL_P, s);
                u64 b = gl_pow(omega % GL_P, s);
                S.step = gl_pow(b, T);
                for (int j = 0; j < 32; ++j) {
                    S.sq[j] = b;
                    b = gl_mul(b, b);
                }
            }
            slots.push_back(S);
        }
        slots.back().end = i + 1;
        CombCol &C = cols[i];
        C.ptr = h_cols[c];
        C.stride = h_strides[c];
        C.planes = h_planes[c];
        C.pad = 0;
        for (int j = 0; j < 3; ++j) {
            C.wa[j] = gl_to_mont(h_wa[3 * c + j] % GL_P);
            C.wb[j] = gl_to_mont(gl_to_mont(h_wb[3 * c + j] % GL_P));
        }
    }
    if (n_cols == 0) {
        for (int j = 0; j < 3; ++j) B2S_CUDA(cudaMemsetAsync(d_out + j * out_stride, 0, sizeof(u64) * N, st));
        return 0;
    }
    CombSlot *d_slots = nullptr;
    CombCol *d_cols = nullptr;
    B2S_CUDA(cudaMallocAsync(&d_slots, sizeof(CombSlot) * slots.size(), st));
    B2S_CUDA(cudaMallocAsync(&d_cols, sizeof(CombCol) * cols.size(), st));
    B2S_CUDA(cudaMemcpyAsync(d_slots, slots.data(), sizeof(CombSlot) * slots.size(), cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemcpyAsync(d_cols, cols.data(), sizeof(CombCol) * cols.size(), cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)((T + 255) / 256);
    if (K == 4)
        comb_kernel<4><<<grid, 256, 0, st>>>(d_slots, (u32)slots.size(), d_cols, T, d_out, out_stride);
    else if (K == 2)
        comb_kernel<2><<<grid, 256, 0, st>>>(d_slots, (u32)slots.size(), d_cols, T, d_out, out_stride);
    else
        comb_kernel<1><<<grid, 256, 0, st>>>(d_slots, (u32)slots.size(), d_cols, T, d_out, out_stride);
    B2S_LAUNCHED();
    cudaFreeAsync(d_slots, st);
    cudaFreeAsync(d_cols, st);
    // the descriptors were staged from pageable host memory: make sure the copies are done before
    // the vectors go out of scope
    B2S_CUDA(cudaStreamSynchronize(st));
    return 0;
}
