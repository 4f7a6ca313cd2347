// quotient.cu -- quotient codewords of the AIR constraints.
//
// Replaces the per-point Python loops of code/table.py:155-286 and
// code/permutation_argument.py:11-20: every constraint is a multivariate polynomial
// (code/multivariate.py: dict exponent-vector -> coefficient) that the reference evaluates at
// every point of the FRI domain with MPolynomial.evaluate (code/multivariate.py:105-116) and
// divides by a zerofier.  Here the host re-factors every dictionary into a Horner program
// (quotient_prog.h) that one thread runs for one constraint at two points; the inverse
// zerofier is computed once per point by a first kernel.  HBM-bound only in name: a point reads
// 8 or 24 B per variable it uses and writes 24 B, the field arithmetic dominates.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "glmont.cuh"
#include "quotient_prog.h"

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

// kinds[v] = 0 as soon as plane 1 or 2 of codeword v holds a non-zero value (the caller presets 1).  The base
// columns of a table reach the quotient step lifted into the extension field (every Table.extend of the
// reference: `[xfield.lift(c) for c in codeword]`, e.g. code/io_table.py:106-107), i.e. with zero upper planes.
__global__ void __launch_bounds__(256) column_kind_kernel(const u64 *__restrict__ cw, u64 N, u32 *kinds) {
    const u32 v = blockIdx.x;
    const u64 *p = cw + ((u64)3 * v + 1) * N;  // planes 1 and 2 are adjacent
    u64 any = 0;
    const u64 step = (u64)gridDim.y * blockDim.x;
    u64 i = (u64)blockIdx.y * blockDim.x + threadIdx.x;
    for (; i + 7 * step < 2 * N; i += 8 * step) {  // eight loads in flight per thread
        u64 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = p[i + k * step];
#pragma unroll
        for (int k = 0; k < 8; ++k) any |= t[k];
    }
    for (; i < 2 * N; i += step) any |= p[i];
    if (__syncthreads_or(any != 0) && threadIdx.x == 0) kinds[v] = 0;
}

// One thread evaluates one constraint at K points and multiplies by the inverse zerofier.
//
// The reference hands every constraint over EXPANDED, a dictionary exponent vector -> coefficient
// (code/multivariate.py), and evaluates it monomial by monomial (:105-116).  The AIR's polynomials are products of
// a few shared forms (instruction selectors x instruction-specific relations, code/processor_table.py:130-217),
// so the host re-factors them (quotient_prog.h): a greedy multivariate Horner scheme
//     P = v * Q + R,   v = the variable that saves the most multiplication work,  Q, R recursively
// which needs 691 base-field multiplications per point for all 47 constraints of the Brainfuck AIR against 3 164
// monomial by monomial (`profiles/microbench/horner_cost.py`).  The arithmetic follows the FIELD each operand lives
// in: base columns reach this step lifted into the extension field with zero upper planes (every Table.extend:
// `[xfield.lift(c) for c in codeword]`), so acc * variable costs 1 (both base-field), 3 (one of them) or 9
// multiplications; the compiler tracks the kinds statically.
//
// The tree is flattened into a program for an accumulator + stack machine, children ordered by their stack need
// (Sethi-Ullman), so the depth is <= log2(monomials) + 1 (3 for this AIR); the stack lives in shared memory as
// [word][point of the thread][thread].  K points per thread share the decoding of every instruction.
//
// Montgomery multiplications by PLAIN codeword values: every factor divides the running value by 2^64, which the
// host has compensated by scaling each monomial's coefficient with 2^(64 * degree) -- a leaf passes through
// exactly `degree` multiplications on its way to the root.
template <int K>
__global__ void __launch_bounds__(Q_THREADS)
    quotient_kernel(const u64 *__restrict__ cw, u64 N, u32 width, u64 shift, const u64 *__restrict__ consts,
                    const u64 *__restrict__ prog, const u32 *__restrict__ prog_off, const u64 *__restrict__ zinv,
                    u64 *__restrict__ out) {
    extern __shared__ u64 q_sm[];
    const u64 i0 = (u64)blockIdx.x * (K * Q_THREADS) + threadIdx.x;  // point k of the thread: i0 + k * Q_THREADS
    const u32 c = blockIdx.y;
    if (K == 1 && i0 >= N) return;  // K > 1 is only launched on domains that fill every thread
    struct Mem {
        u64 *mine;
        const u64 *here[K], *next[K];
        u64 N;
        u32 width;
        __device__ __forceinline__ u64 var(u32 v, int j, int k) const {
            return ((v >= width ? next[k] : here[k]) + ((u64)3 * v + j) * N)[0];
        }
        __device__ __forceinline__ u64 get(u32 w, int k) const { return mine[(w * K + k) * Q_THREADS]; }
        __device__ __forceinline__ void put(u32 w, int k, u64 x) { mine[(w * K + k) * Q_THREADS] = x; }
    } mem;
    mem.mine = q_sm + threadIdx.x;
    mem.N = N;
    mem.width = width;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const u64 i = i0 + (u64)k * Q_THREADS;
        u64 inext = i + shift;
        if (inext >= N) inext -= N;
        mem.here[k] = cw + i;
        mem.next[k] = cw + inext - (u64)3 * width * N;
    }
    xfe acc[K];
    q_run<K>(prog + prog_off[c], consts, mem, acc);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const u64 i = i0 + (u64)k * Q_THREADS;
        const u64 zm = zinv[i];
        u64 *o = out + (u64)3 * c * N + i;
        o[0] = lcanon(mont_mul(acc[k].c[0], zm));
        o[N] = lcanon(mont_mul(acc[k].c[1], zm));
        o[2 * N] = lcanon(mont_mul(acc[k].c[2], zm));
    }
}

}  // namespace

extern "C" int b2s_quotients(const uint64_t *d_cw, uint64_t N, uint32_t width, uint64_t shift, uint32_t n_constraints,
                             const uint32_t *h_mono_off, const uint64_t *h_coeffs, const uint32_t *h_factors,
                             uint32_t max_factors, uint32_t zerofier_kind, uint64_t height, uint64_t omicron_inv,
                             uint64_t offset, uint64_t omega, uint64_t *d_out, int *h_zero_flag,
                             const uint8_t *h_base_columns, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || (N & (N - 1)) || width == 0 || zerofier_kind < 1 || zerofier_kind > 3) {
        b2s_set_error("quotients: bad arguments (N %llu, width %u, shift %llu, zerofier %u)", (unsigned long long)N, width,
                      (unsigned long long)shift, zerofier_kind);
        return B2S_ERR_ARG;
    }
    shift &= N - 1;  // rows are taken modulo N: a table of height 1 has unit distance N (code/table.py:37-40)
    if (h_zero_flag) *h_zero_flag = 0;
    if (n_constraints == 0) return 0;
    const u32 n_mono = h_mono_off[n_constraints];
    for (u32 m = 0; m < n_mono; ++m)
        for (u32 f = 0; f < max_factors; ++f) {
            const u32 fac = h_factors[m * max_factors + f];
            if ((fac & 0xFF) && (fac >> 8) >= 2 * width) {
                b2s_set_error("quotients: monomial %u uses variable %u of %u", m, fac >> 8, 2 * width);
                return B2S_ERR_ARG;
            }
        }
    // which codewords are lifted base-field columns: the caller's word for it (it zeroed the upper planes itself),
    // else an exact scan of the columns
    std::vector<u32> kinds(width, 0);
    if (h_base_columns) {
        for (u32 v = 0; v < width; ++v) kinds[v] = h_base_columns[v] != 0;
    } else {
        u32 *d_kinds = nullptr;
        B2S_CUDA(cudaMallocAsync(&d_kinds, sizeof(u32) * width, st));
        B2S_CUDA(cudaMemsetAsync(d_kinds, 1, sizeof(u32) * width, st));  // non-zero = base-field until proven otherwise
        const unsigned per_col = (unsigned)std::min<u64>(296, (2 * N + 2047) / 2048);
        column_kind_kernel<<<dim3(width, per_col), 256, 0, st>>>(d_cw, N, d_kinds);
        B2S_LAUNCHED();
        B2S_CUDA(cudaMemcpyAsync(kinds.data(), d_kinds, sizeof(u32) * width, cudaMemcpyDeviceToHost, st));
        B2S_CUDA(cudaStreamSynchronize(st));
        cudaFreeAsync(d_kinds, st);
    }
    // compile (quotient_prog.h): expanded monomials -> greedy Horner tree -> stack program; one blob per call:
    //     constants (3 words each) | zero flag, pad | programs (64-bit instructions) | prog_off
    std::vector<u64> consts, code;
    std::vector<u32> prog_off;
    u32 max_need = 0;
    char why[160];
    if (q_compile(width, n_constraints, h_mono_off, h_coeffs, h_factors, max_factors, kinds, consts, code, prog_off, max_need,
                  why, sizeof(why))) {
        b2s_set_error("quotients: %s", why);
        return B2S_ERR_ARG;
    }
    const size_t n_coef = consts.size();
    std::vector<u64> blob(n_coef + 1 + code.size() + (n_constraints + 2) / 2, 0);
    memcpy(blob.data(), consts.data(), sizeof(u64) * n_coef);
    memcpy(blob.data() + n_coef + 1, code.data(), sizeof(u64) * code.size());
    memcpy(blob.data() + n_coef + 1 + code.size(), prog_off.data(), sizeof(u32) * (n_constraints + 1));
    u64 *d_blob = nullptr, *d_zinv = nullptr;
    B2S_CUDA(cudaMallocAsync(&d_blob, sizeof(u64) * blob.size(), st));
    B2S_CUDA(cudaMallocAsync(&d_zinv, sizeof(u64) * N, st));
    B2S_CUDA(cudaMemcpyAsync(d_blob, blob.data(), sizeof(u64) * blob.size(), cudaMemcpyHostToDevice, st));
    const u64 *d_coef = d_blob;
    int *d_flag = reinterpret_cast<int *>(d_blob + n_coef);
    const u64 *d_code = d_blob + n_coef + 1;
    const u32 *d_poff = reinterpret_cast<const u32 *>(d_code + code.size());
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
    const int K = N >= ((u64)1 << 18) ? 32 : N >= 8 * 256 ? 8 : 1;  // points per field inversion
    const u64 T = N / K;
    Z.step = gl_pow(omega % GL_P, T);
    Z.step_h = gl_pow(Z.step, height);
    if (K == 32)
        zerofier_kernel<32><<<(unsigned)((T + 63) / 64), 64, 0, st>>>(Z, T, d_zinv, d_flag);
    else if (K == 8)
        zerofier_kernel<8><<<(unsigned)((T + 255) / 256), 256, 0, st>>>(Z, T, d_zinv, d_flag);
    else
        zerofier_kernel<1><<<(unsigned)((T + 255) / 256), 256, 0, st>>>(Z, T, d_zinv, d_flag);
    B2S_LAUNCHED();
    // points per thread: 2 once the domain fills the machine twice over (env B2S_Q_K overrides: 1 or 2; 4 measured no faster)
    static const int k_env = getenv("B2S_Q_K") ? atoi(getenv("B2S_Q_K")) : 0;
    int QK = k_env ? k_env : (N >= ((u64)1 << 16) ? 2 : 1);
    if (N < (u64)QK * Q_THREADS || (QK != 1 && QK != 2)) QK = 1;
    const size_t smem = sizeof(u64) * Q_THREADS * 3 * (size_t)std::max<u32>(max_need, 1) * QK;
    const dim3 grid((unsigned)((N + (u64)QK * Q_THREADS - 1) / ((u64)QK * Q_THREADS)), n_constraints);
    auto launch = [&](auto kernel) -> int {
        if (smem > 48 * 1024)  // per device and cheap: no caching across calls
            B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, Q_THREADS, smem, st>>>(d_cw, N, width, shift, d_coef, d_code, d_poff, d_zinv, d_out);
        B2S_LAUNCHED();
        return 0;
    };
    if (int rc = QK == 2 ? launch(quotient_kernel<2>) : launch(quotient_kernel<1>)) return rc;
    if (!h_zero_flag) {  // the caller has ruled a vanishing zerofier out: nothing to read back, the call stays asynchronous
        cudaFreeAsync(d_blob, st);
        cudaFreeAsync(d_zinv, st);
        return 0;
    }
    int flag = 0;
    B2S_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    cudaFreeAsync(d_blob, st);
    cudaFreeAsync(d_zinv, st);
    B2S_CUDA(cudaStreamSynchronize(st));
    *h_zero_flag = flag;
    return 0;
}
