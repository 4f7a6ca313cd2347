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
#include <string.h>

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
#define Q_BASE 0x80000000u  // op / hot word: the variable's codeword is a lifted base-field column (planes 1, 2 zero)

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

// One thread evaluates one (constraint, point) pair:  sum_m coeff_m prod_f var_f^e_f, times the inverse zerofier.
// The arithmetic follows the FIELD each operand lives in: most variables of the Brainfuck AIR are base-field
// columns, so a monomial is evaluated as  (product of its base-field factors: 1 multiplication each)  x  coefficient
// (1 or 3)  x  extension-field factors (9 each).  The processor table's transition constraints drop from 11 889 to
// 2 669 base-field multiplications per point.  The host COMPILES every constraint into a word stream that the
// kernel walks without tests (r02h profile: 45 % of the executed instructions of the first version decoded
// factor slots):
//     per monomial   header = n_factors | hot_exponent << 16 | coefficient_is_extension << 24
//                    n_factors words: variable << 8 | exponent | Q_BASE for a base-field column   (exponent >= 1),
//                    base-field factors first
// in Horner order of the constraint's hot variable (below); coefficients in the same order in `coeffs`.
//
// hot[c] = variable | (max cached exponent << 16) | Q_BASE, or ~0.  The processor table's instruction selectors carry
// one (base-field) variable to every power up to 8 in dozens of monomials (code/processor_table.py:130-217).  Its
// powers are built once per thread, and the monomials are sorted by descending exponent of that variable, so the
// sum is evaluated as a polynomial in it by Horner's rule:
//     sum_m c_m h^(e_m) rest_m  =  (...((S_8) h + S_7) h + ...) h + S_0,   S_e = sum of c_m rest_m with e_m = e.
//
// Montgomery multiplications by PLAIN codeword values: every factor divides the running product by 2^64, which the
// host has compensated by scaling the monomial's coefficient with 2^(64 * degree).  Cached powers keep that
// bookkeeping: c_1 = x, c_(k+1) = c_k * x * 2^-64, so acc * c_e * 2^-64 equals e successive multiplications by x.
__global__ void __launch_bounds__(Q_THREADS)
    quotient_kernel(const u64 *__restrict__ cw, u64 N, u32 width, u64 shift, const u32 *__restrict__ mono_off,
                    const u64 *__restrict__ coeffs, const u32 *__restrict__ prog, const u32 *__restrict__ prog_off,
                    const u32 *__restrict__ hot, const u64 *__restrict__ zinv, u64 *__restrict__ out) {
    __shared__ u64 pw[Q_HOT_MAX * 3 * Q_THREADS];  // [exponent - 1][coefficient][thread]
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u32 c = blockIdx.y;
    if (i >= N) return;
    u64 inext = i + shift;
    if (inext >= N) inext -= N;
    const u64 *here = cw + i, *next = cw + inext - (u64)3 * width * N;
    auto plane0 = [&](u32 v) { return (v >= width ? next : here) + (u64)3 * v * N; };
    auto load = [&](u32 v) {
        const u64 *p = plane0(v);
        return xfe{{p[0], p[N], p[2 * N]}};
    };
    const u32 hv = hot[c];
    const u32 hvar = hv & 0xFFFF, hmax = hv == 0xFFFFFFFFu ? 0 : (hv >> 16) & 0xFF;
    const bool hbase = hmax && (hv & Q_BASE);
    u64 *mine = pw + threadIdx.x;
    if (hmax) {
        if (hbase) {
            const u64 x = *plane0(hvar);
            u64 cur = x;
            for (u32 e = 1;; ++e) {
                mine[(e - 1) * 3 * Q_THREADS] = cur;
                if (e == hmax) break;
                cur = mont_mul(cur, x);
            }
        } else {
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
    }
    // acc * h^e (2^-64 bookkeeping as above)
    auto times_hot = [&](xfe &acc, u32 e) {
        const u64 *q = mine + (e - 1) * 3 * Q_THREADS;
        if (hbase) {
            const u64 h = q[0];
#pragma unroll
            for (int j = 0; j < 3; ++j) acc.c[j] = mont_mul(acc.c[j], h);
        } else {
            acc = x_mul_mont(acc, xfe{{q[0], q[Q_THREADS], q[2 * Q_THREADS]}});
        }
    };
    xfe acc = {{0, 0, 0}};
    u32 level = 0;  // exponent of the hot variable that acc still has to be multiplied by
    const u32 *pc = prog + prog_off[c];
    const u32 m0 = mono_off[c], m1 = mono_off[c + 1];
    for (u32 m = m0; m < m1; ++m) {
        const u32 hdr = *pc++;
        const u32 nops = hdr & 0xFFFF, eh = (hdr >> 16) & 0xFF;
        const bool cext = (hdr >> 24) != 0;
        if (m == m0) {
            level = eh;
        } else if (eh < level) {  // monomials arrive by descending eh
            times_hot(acc, level - eh);
            level = eh;
        }
        const u64 c0 = coeffs[3 * m];
        u64 c1 = 0, c2 = 0;
        if (cext) {
            c1 = coeffs[3 * m + 1];
            c2 = coeffs[3 * m + 2];
        }
        // b = product of the leading base-field factors; `prod` takes over at the first extension-field factor
        u64 b = 0;
        bool have_b = false, ext = false;
        xfe prod = {{0, 0, 0}};
        for (u32 f = 0; f < nops; ++f) {
            const u32 op = *pc++;
            const u32 v = (op >> 8) & 0xFFFF;
            u32 e = op & 0xFF;
            if (op & Q_BASE) {
                const u64 x = *plane0(v);
                if (ext) {
                    for (; e; --e)
#pragma unroll
                        for (int j = 0; j < 3; ++j) prod.c[j] = mont_mul(prod.c[j], x);
                } else {
                    if (!have_b) {
                        b = x;
                        have_b = true;
                        --e;
                    }
                    for (; e; --e) b = mont_mul(b, x);
                }
            } else {
                const xfe x = load(v);
                if (!ext) {
                    ext = true;
                    if (!cext) {  // base-field scalar times the first extension-field factor: 3 multiplications
                        const u64 sc = have_b ? mont_mul(c0, b) : c0;
#pragma unroll
                        for (int j = 0; j < 3; ++j) prod.c[j] = mont_mul(x.c[j], sc);
                    } else {
                        prod = have_b ? xfe{{mont_mul(c0, b), mont_mul(c1, b), mont_mul(c2, b)}} : xfe{{c0, c1, c2}};
                        prod = x_mul_mont(prod, x);
                    }
                    --e;
                }
                for (; e; --e) prod = x_mul_mont(prod, x);
            }
        }
        if (ext) {
#pragma unroll
            for (int j = 0; j < 3; ++j) acc.c[j] = ladd(acc.c[j], lcanon(prod.c[j]));
        } else if (have_b) {
            acc.c[0] = ladd(acc.c[0], mont_mul(c0, b));  // mont_mul results are canonical
            if (cext) {
                acc.c[1] = ladd(acc.c[1], mont_mul(c1, b));
                acc.c[2] = ladd(acc.c[2], mont_mul(c2, b));
            }
        } else {
            acc.c[0] = ladd(acc.c[0], c0);  // host-scaled coefficients are canonical
            if (cext) {
                acc.c[1] = ladd(acc.c[1], c1);
                acc.c[2] = ladd(acc.c[2], c2);
            }
        }
    }
    if (level) times_hot(acc, level);
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
                             uint64_t offset, uint64_t omega, uint64_t *d_out, int *h_zero_flag,
                             const uint8_t *h_base_columns, void *stream) {
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
    auto is_base = [&](u32 v) { return kinds[v >= width ? v - width : v] != 0; };
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
            hot[c] = best | (maxe[best] << 16) | (is_base(best) ? Q_BASE : 0);
    }
    // ... its monomials sorted by descending exponent of that variable (the kernel's Horner order), and every
    // coefficient times 2^(64 * total degree) (see quotient_kernel)
    auto hot_exp = [&](u32 c, u32 m) -> u32 {
        if (hot[c] == 0xFFFFFFFFu) return 0;
        const u32 hvar = hot[c] & 0xFFFF, hmax = (hot[c] >> 16) & 0xFF;
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
    std::vector<u32> code, prog_off(n_constraints + 1, 0);
    code.reserve((size_t)n_mono * (mf + 1) + 1);
    for (u32 c = 0; c < n_constraints; ++c) {
        prog_off[c] = (u32)code.size();
        const bool has_hot = hot[c] != 0xFFFFFFFFu;
        const u32 hvar = hot[c] & 0xFFFF, hmax = (hot[c] >> 16) & 0xFF;
        for (u32 k = h_mono_off[c]; k < h_mono_off[c + 1]; ++k) {
            const u32 m = order[k];
            u64 degree = 0;
            u32 eh = 0;
            std::vector<u32> ops;
            for (u32 f = 0; f < max_factors; ++f) {
                const u32 fac = h_factors[m * max_factors + f], e = fac & 0xFF, v = fac >> 8;
                if (e == 0) continue;
                degree += e;
                if (has_hot && v == hvar && e <= hmax && eh == 0)
                    eh = e;  // applied to the whole group by the Horner step (a repeated factor stays generic)
                else
                    ops.push_back((v << 8) | e | (is_base(v) ? Q_BASE : 0));
            }
            std::stable_sort(ops.begin(), ops.end(), [](u32 a, u32 b) { return (a & Q_BASE) > (b & Q_BASE); });  // base-field factors lead
            const u64 r = gl_pow(GL_EPS, degree);  // 2^64 = EPS (mod p)
            for (int j = 0; j < 3; ++j) scaled[3 * k + j] = gl_mul(h_coeffs[3 * m + j] % GL_P, r);
            const u32 cext = (scaled[3 * k + 1] | scaled[3 * k + 2]) ? 1 : 0;
            code.push_back((u32)ops.size() | (eh << 16) | (cext << 24));
            code.insert(code.end(), ops.begin(), ops.end());
        }
    }
    prog_off[n_constraints] = (u32)code.size();
    code.push_back(0);
    // The program, a few KB: ONE allocation and ONE upload per call (five separate ones cost more host time than the
    // small tables' kernels take):  coefficients | zero flag, pad | mono_off | prog_off | hot | code
    const size_t n_coef = 3 * (size_t)n_mono + 1;
    std::vector<u64> blob(n_coef + 1 + ((size_t)2 * (n_constraints + 1) + n_constraints + code.size() + 1) / 2 + 1, 0);
    memcpy(blob.data(), scaled.data(), sizeof(u64) * 3 * (size_t)n_mono);
    u32 *words = reinterpret_cast<u32 *>(blob.data() + n_coef + 1);
    memcpy(words, h_mono_off, sizeof(u32) * (n_constraints + 1));
    memcpy(words + (n_constraints + 1), prog_off.data(), sizeof(u32) * (n_constraints + 1));
    memcpy(words + 2 * (n_constraints + 1), hot.data(), sizeof(u32) * n_constraints);
    memcpy(words + 2 * (n_constraints + 1) + n_constraints, code.data(), sizeof(u32) * code.size());
    u64 *d_blob = nullptr, *d_zinv = nullptr;
    B2S_CUDA(cudaMallocAsync(&d_blob, sizeof(u64) * blob.size(), st));
    B2S_CUDA(cudaMallocAsync(&d_zinv, sizeof(u64) * N, st));
    B2S_CUDA(cudaMemcpyAsync(d_blob, blob.data(), sizeof(u64) * blob.size(), cudaMemcpyHostToDevice, st));
    const u64 *d_coef = d_blob;
    int *d_flag = reinterpret_cast<int *>(d_blob + n_coef);
    const u32 *d_words = reinterpret_cast<const u32 *>(d_blob + n_coef + 1);
    const u32 *d_off = d_words, *d_poff = d_words + (n_constraints + 1), *d_hot = d_words + 2 * (n_constraints + 1);
    const u32 *d_fac = d_hot + n_constraints;
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
    quotient_kernel<<<dim3((unsigned)((N + Q_THREADS - 1) / Q_THREADS), n_constraints), Q_THREADS, 0, st>>>(
        d_cw, N, width, shift, d_off, d_coef, d_fac, d_poff, d_hot, d_zinv, d_out);
    B2S_LAUNCHED();
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
