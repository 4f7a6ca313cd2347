// quotient.cu -- quotient codewords of the AIR constraints.
//
// Replaces the per-point Python loops of code/table.py:155-286 and
// code/permutation_argument.py:11-20: every constraint is a multivariate polynomial
// (code/multivariate.py: dict exponent-vector -> coefficient) that the reference evaluates at
// every point of the FRI domain with MPolynomial.evaluate (code/multivariate.py:105-116) and
// divides by a zerofier.  Here the host flattens the dictionaries into a monomial program and
// one thread evaluates one (constraint, point) pair in extension-field arithmetic; the inverse
// zerofier is computed once per point by a first kernel.  HBM-bound only in name: a point reads
// 24 B per variable it uses and writes 24 B, the monomial arithmetic dominates.
#include <algorithm>
#include <vector>

#include "common.h"
#include "glmont.cuh"

namespace {

struct ZeroParams {
    u64 w_sq[32];  // omega^(2^b)
    u64 offset, omicron_inv, height;
    u64 step;      // omega^T: from one point of a thread to its next
    u64 step_h;    // omega^(T * height)
    u32 kind;
};

// zinv[i] = 2^64 * inverse zerofier at x_i = offset * omega^i; *flag = 1 if a zerofier vanishes.
// A thread owns K points i = t + k*T and inverts their K zerofier values with ONE field inversion
// (Montgomery's trick: prefix products, invert the last, walk back) -- the inversion is ~100
// multiplications, the rest 3 per point.  A vanishing value poisons the thread's K results, but
// then the flag is set and the caller raises the reference's assertion anyway.
template <int K>
__global__ void __launch_bounds__(256) zerofier_kernel(const __grid_constant__ ZeroParams Z, u64 T, u64 *zinv, int *flag) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    if (Z.kind == B2S_ZEROFIER_TRANSITION && Z.height == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) zinv[t + k * T] = 0;  // x^0 - 1 = 0 is used as is (code/table.py:196-199)
        return;
    }
    u64 x = gl_mul(Z.offset, gl_pow_sq(Z.w_sq, t));
    u64 xh = Z.kind == B2S_ZEROFIER_TRANSITION ? gl_pow(x, Z.height) : 0;
    u64 z[K], mul[K], pre[K];
    bool zero = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (Z.kind == B2S_ZEROFIER_BOUNDARY) {
            z[k] = gl_sub(x, 1);
            mul[k] = 1;
        } else if (Z.kind == B2S_ZEROFIER_TRANSITION) {
            z[k] = gl_sub(xh, 1);
            mul[k] = gl_sub(x, Z.omicron_inv);
            xh = gl_mul(xh, Z.step_h);
        } else {
            z[k] = gl_sub(x, Z.omicron_inv);
            mul[k] = 1;
        }
        zero |= z[k] == 0;
        pre[k] = k ? gl_mul(pre[k - 1], z[k]) : z[k];
        x = gl_mul(x, Z.step);
    }
    if (zero) *flag = 1;
    u64 inv = gl_inv(pre[K - 1]);
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        const u64 r = k ? gl_mul(inv, pre[k - 1]) : inv;
        if (k) inv = gl_mul(inv, z[k]);
        // Montgomery form: the final multiplication of quotient_kernel then needs no extra reduction
        zinv[t + k * T] = gl_to_mont(Z.kind == B2S_ZEROFIER_TRANSITION ? gl_mul(r, mul[k]) : r);
    }
}

#define Q_HOT_MAX 8      // cached powers of a constraint's most-exponentiated variable
#define Q_THREADS 128

// hot[c] = variable | (max cached exponent << 16) or ~0.  The processor table's instruction selectors carry
// one variable to every power up to 8 in dozens of monomials (code/processor_table.py:130-217): 1 000 of the
// 1 321 extension-field multiplications per point of its transition constraints.  Its powers are built once
// per thread, and the host has sorted the constraint's monomials by descending exponent of that variable, so
// the sum is evaluated as a polynomial in it by Horner's rule:
//     sum_m c_m h^(e_m) rest_m  =  (...((S_8) h + S_7) h + ...) h + S_0,   S_e = sum of c_m rest_m with e_m = e.
__global__ void __launch_bounds__(Q_THREADS)
    quotient_kernel(const u64 *__restrict__ cw, u64 N, u32 width, u64 shift, const u32 *__restrict__ mono_off,
                    const u64 *__restrict__ coeffs, const u32 *__restrict__ factors, u32 max_factors,
                    const u32 *__restrict__ hot, const u64 *__restrict__ zinv, u64 *__restrict__ out) {
    __shared__ u64 pw[Q_HOT_MAX * 3 * Q_THREADS];  // [exponent - 1][coefficient][thread]
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 c = blockIdx.y;
    if (i >= N) return;
    u64 inext = i + shift;
    if (inext >= N) inext -= N;
    auto load = [&](u32 v) {
        u64 at = i;
        if (v >= width) {
            v -= width;
            at = inext;
        }
        const u64 *p = cw + (u64)3 * v * N + at;
        return xfe{{p[0], p[N], p[2 * N]}};
    };
    // Montgomery multiplications by PLAIN codeword values: every factor divides the running product by
    // 2^64, which the host has compensated by scaling the monomial's coefficient with 2^(64 * degree).
    // Cached powers keep that bookkeeping: c_1 = x, c_(k+1) = c_k * x * 2^-64, so acc * c_e * 2^-64
    // equals e successive multiplications by x.
    const u32 hv = hot[c];
    const u32 hvar = hv & 0xFFFF, hmax = hv == 0xFFFFFFFFu ? 0 : hv >> 16;
    u64 *mine = pw + threadIdx.x;
    auto power = [&](u32 e) {
        const u64 *q = mine + (e - 1) * 3 * Q_THREADS;
        return xfe{{q[0], q[Q_THREADS], q[2 * Q_THREADS]}};
    };
    if (hmax) {
        const xfe x = load(hvar);
        xfe cur = x;
        for (u32 e = 1;; ++e) {
#pragma unroll
            for (int j = 0; j < 3; ++j) mine[((e - 1) * 3 + j) * Q_THREADS] = cur.c[j];
            if (e == hmax) break;
            cur = x_mul_mont(cur, x);
#pragma unroll
            for (int j = 0; j < 3; ++j) cur.c[j] = lcanon(cur.c[j]);
        }
    }
    xfe acc = {{0, 0, 0}};
    u32 level = 0;  // exponent of the hot variable that acc still has to be multiplied by
    for (u32 m = mono_off[c]; m < mono_off[c + 1]; ++m) {
        xfe prod = {{coeffs[3 * m], coeffs[3 * m + 1], coeffs[3 * m + 2]}};
        u32 eh = 0;
        for (u32 f = 0; f < max_factors; ++f) {
            const u32 fac = factors[m * max_factors + f];
            const u32 e = fac & 0xFF;
            if (e == 0) continue;
            const u32 v = fac >> 8;
            if (v == hvar && e <= hmax && eh == 0) {  // (a repeated factor of the same variable stays generic)
                eh = e;  // applied to the whole group by the Horner step below
            } else {
                const xfe x = load(v);
                for (u32 k = 0; k < e; ++k) prod = x_mul_mont(prod, x);
            }
        }
        if (m == mono_off[c]) {
            level = eh;
        } else if (eh < level) {  // monomials arrive by descending eh
            acc = x_mul_mont(acc, power(level - eh));
            level = eh;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) acc.c[j] = ladd(acc.c[j], lcanon(prod.c[j]));
    }
    if (level) acc = x_mul_mont(acc, power(level));
    const u64 zm = zinv[i];
    u64 *o = out + (u64)3 * c * N + i;
    o[0] = lcanon(mont_mul(acc.c[0], zm));
    o[N] = lcanon(mont_mul(acc.c[1], zm));
    o[2 * N] = lcanon(mont_mul(acc.c[2], zm));
}

}  // namespace

extern "C" int b2s_quotients(const uint64_t *d_cw, uint64_t N, uint32_t width, uint64_t shift, uint32_t n_constraints,
                             const uint32_t *h_mono_off, const uint64_t *h_coeffs, const uint32_t *h_factors,
                             uint32_t max_factors, uint32_t zerofier_kind, uint64_t height, uint64_t omicron_inv,
                             uint64_t offset, uint64_t omega, uint64_t *d_out, int *h_zero_flag, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || (N & (N - 1)) || shift >= N || width == 0 || zerofier_kind < 1 || zerofier_kind > 3) {
        b2s_set_error("quotients: bad arguments (N %llu, width %u, shift %llu, zerofier %u)", (unsigned long long)N, width,
                      (unsigned long long)shift, zerofier_kind);
        return B2S_ERR_ARG;
    }
    if (h_zero_flag) *h_zero_flag = 0;
    if (n_constraints == 0) return 0;
    const u32 n_mono = h_mono_off[n_constraints];
    const u32 mf = max_factors ? max_factors : 1;
    for (u32 m = 0; m < n_mono; ++m)
        for (u32 f = 0; f < max_factors; ++f) {
            const u32 fac = h_factors[m * max_factors + f];
            if ((fac & 0xFF) && (fac >> 8) >= 2 * width) {
                b2s_set_error("quotients: monomial %u uses variable %u of %u", m, fac >> 8, 2 * width);
                return B2S_ERR_ARG;
            }
        }
    // the program: a few KB, uploaded per call
    u32 *d_off = nullptr, *d_fac = nullptr, *d_hot = nullptr;
    u64 *d_coef = nullptr, *d_zinv = nullptr;
    int *d_flag = nullptr;
    B2S_CUDA(cudaMallocAsync(&d_off, sizeof(u32) * (n_constraints + 1), st));
    B2S_CUDA(cudaMallocAsync(&d_fac, sizeof(u32) * ((size_t)n_mono * mf + 1), st));
    B2S_CUDA(cudaMallocAsync(&d_coef, sizeof(u64) * (3 * (size_t)n_mono + 1), st));
    B2S_CUDA(cudaMallocAsync(&d_zinv, sizeof(u64) * N, st));
    B2S_CUDA(cudaMallocAsync(&d_flag, sizeof(int), st));
    B2S_CUDA(cudaMemcpyAsync(d_off, h_mono_off, sizeof(u32) * (n_constraints + 1), cudaMemcpyHostToDevice, st));
    // per constraint: the variable whose powers are worth caching and factoring out (most multiplications saved) ...
    std::vector<u32> hot(n_constraints, 0xFFFFFFFFu);
    for (u32 c = 0; c < n_constraints; ++c) {
        std::vector<u64> saved(2 * (size_t)width, 0);
        std::vector<u32> maxe(2 * (size_t)width, 0);
        for (u32 m = h_mono_off[c]; m < h_mono_off[c + 1]; ++m)
            for (u32 f = 0; f < max_factors; ++f) {
                const u32 fac = h_factors[m * max_factors + f], e = fac & 0xFF, v = fac >> 8;
                if (e >= 1 && e <= Q_HOT_MAX) {
                    saved[v] += e;  // Horner applies the variable once per exponent level, not per monomial
                    maxe[v] = std::max(maxe[v], e);
                }
            }
        u32 best = 0;
        for (u32 v = 1; v < 2 * width; ++v)
            if (saved[v] > saved[best]) best = v;
        if (maxe[best] >= 2 && saved[best] > 2 * (u64)maxe[best] && 2 * width <= 0xFFFF)
            hot[c] = best | (maxe[best] << 16);
    }
    // ... its monomials sorted by descending exponent of that variable (the kernel's Horner order), and every
    // coefficient times 2^(64 * total degree) (see quotient_kernel)
    auto hot_exp = [&](u32 c, u32 m) -> u32 {
        if (hot[c] == 0xFFFFFFFFu) return 0;
        const u32 hvar = hot[c] & 0xFFFF, hmax = hot[c] >> 16;
        for (u32 f = 0; f < max_factors; ++f) {
            const u32 fac = h_factors[m * max_factors + f], e = fac & 0xFF;
            if (e && (fac >> 8) == hvar && e <= hmax) return e;
        }
        return 0;
    };
    std::vector<u32> order(n_mono);
    for (u32 m = 0; m < n_mono; ++m) order[m] = m;
    for (u32 c = 0; c < n_constraints; ++c)
        std::stable_sort(order.begin() + h_mono_off[c], order.begin() + h_mono_off[c + 1],
                         [&](u32 a, u32 b) { return hot_exp(c, a) > hot_exp(c, b); });
    std::vector<u64> scaled(3 * (size_t)n_mono + 1);
    std::vector<u32> facs((size_t)n_mono * mf + 1, 0);
    for (u32 k = 0; k < n_mono; ++k) {
        const u32 m = order[k];
        u64 degree = 0;
        for (u32 f = 0; f < max_factors; ++f) {
            facs[(size_t)k * max_factors + f] = h_factors[m * max_factors + f];
            degree += h_factors[m * max_factors + f] & 0xFF;
        }
        const u64 r = gl_pow(GL_EPS, degree);  // 2^64 = EPS (mod p)
        for (int j = 0; j < 3; ++j) scaled[3 * k + j] = gl_mul(h_coeffs[3 * m + j] % GL_P, r);
    }
    if (n_mono) {
        B2S_CUDA(cudaMemcpyAsync(d_fac, facs.data(), sizeof(u32) * (size_t)n_mono * max_factors, cudaMemcpyHostToDevice, st));
        B2S_CUDA(cudaMemcpyAsync(d_coef, scaled.data(), sizeof(u64) * 3 * (size_t)n_mono, cudaMemcpyHostToDevice, st));
    }
    B2S_CUDA(cudaMallocAsync(&d_hot, sizeof(u32) * n_constraints, st));
    B2S_CUDA(cudaMemcpyAsync(d_hot, hot.data(), sizeof(u32) * n_constraints, cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    ZeroParams Z;
    u64 sq = omega;
    for (int b = 0; b < 32; ++b) {
        Z.w_sq[b] = sq;
        sq = gl_mul(sq, sq);
    }
    Z.offset = offset;
    Z.omicron_inv = omicron_inv;
    Z.height = height;
    Z.kind = zerofier_kind;
    const int K = N >= 8 * 256 ? 8 : 1;
    const u64 T = N / K;
    Z.step = gl_pow(omega % GL_P, T);
    Z.step_h = gl_pow(Z.step, height);
    if (K == 8)
        zerofier_kernel<8><<<(unsigned)((T + 255) / 256), 256, 0, st>>>(Z, T, d_zinv, d_flag);
    else
        zerofier_kernel<1><<<(unsigned)((T + 255) / 256), 256, 0, st>>>(Z, T, d_zinv, d_flag);
    B2S_LAUNCHED();
    quotient_kernel<<<dim3((unsigned)((N + Q_THREADS - 1) / Q_THREADS), n_constraints), Q_THREADS, 0, st>>>(
        d_cw, N, width, shift, d_off, d_coef, d_fac, max_factors, d_hot, d_zinv, d_out);
    B2S_LAUNCHED();
    int flag = 0;
    B2S_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    cudaFreeAsync(d_off, st);
    cudaFreeAsync(d_fac, st);
    cudaFreeAsync(d_hot, st);
    cudaFreeAsync(d_coef, st);
    cudaFreeAsync(d_zinv, st);
    cudaFreeAsync(d_flag, st);
    B2S_CUDA(cudaStreamSynchronize(st));
    if (h_zero_flag) *h_zero_flag = flag;
    return 0;
}
