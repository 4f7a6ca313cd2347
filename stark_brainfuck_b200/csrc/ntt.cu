// ntt.cu -- batched radix-2^q number-theoretic transforms over F_p (p = 2^64 - 2^32 + 1).
//
// Replaces code/ntt.py:4-42 (ntt/intt) and the scale+pad+ntt wrappers of code/ntt.py:164-174
// and code/fri.py:26-44.  Natural order in, natural order out:
//     out[k] = sum_j in[j] * omega^(j*k)
//
// Schedule (not the reference's recursion): a transform of length n = 2^log_n is split
// into 1-3 PASSES over global memory (Cooley-Tukey on the index digits, most significant
// input digit first).  A pass handles sub-transforms of length L interleaved with stride I
// and removes one digit of radix R = R1*R2 <= 1024:
//
//   tile   : R rows x T contiguous columns; row r of column q lives at in[r*(L/R)*I + q]
//   phase 1: each thread gathers R1 rows straight from global memory into registers
//            (fused: zero padding beyond n_in and the coset scale offset^j), runs an R1-point
//            NTT in registers, multiplies by omega_R^(j'*k) and scatters to shared memory
//   phase 2: each thread gathers R2 rows from shared memory, runs an R2-point NTT in
//            registers, applies the inter-pass twiddle omega_L^(j'*k) (or, in the last
//            pass, the fused n^-1 * offset^-k scale of the inverse coset transform) and
//            stores straight to global memory at ((j'*R + k)*I + c)
//
// so every element crosses shared memory once per pass and global memory twice per pass;
// with the vector and its scratch resident in the 126 MB L2, HBM sees the compulsory
// read + write only.  The omega_R power table (R entries) is staged into shared memory by
// one TMA bulk copy (cp.async.bulk + mbarrier) that overlaps the phase-1 global gathers.
#include <map>
#include <mutex>
#include <vector>

#include "common.h"

namespace {

enum : u32 { F_FIRST_SCALE = 1, F_LAST_SCALE = 2, F_CFAST = 4, F_FIRST = 8 };

struct PassParams {
    const u64 *in;
    u64 *out;
    u64 in_stride, out_stride;  // plane strides in elements
    const u64 *tw;              // omega_R^e, e < R (device memory)
    u64 out_mul;                // n^-1 for the inverse transform
    u32 log_n, log_L, log_I;
    u32 n_in;
    u32 flags;
    u64 w_sq[32];  // omega^(2^b)
    u64 s_sq[32];  // scale^(2^b): coset offset (forward, first pass) or its inverse (inverse, last pass)
};

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

template <int Q>
__device__ __forceinline__ constexpr int brev(int i) {
    int r = 0;
    for (int b = 0; b < Q; ++b) r |= ((i >> b) & 1) << (Q - 1 - b);
    return r;
}

// 2^Q-point decimation-in-frequency NTT on registers.  tw[i * tstride] = w^i for the
// primitive 2^Q-th root w.  On return x[i] holds X[brev(i)].
template <int Q>
__device__ __forceinline__ void reg_ntt(u64 (&x)[1 << Q], const u64 *tw, int tstride) {
#pragma unroll
    for (int s = Q - 1; s >= 0; --s) {
        const int half = 1 << s;
#pragma unroll
        for (int i = 0; i < (1 << Q); ++i) {
            if (i & half) continue;
            const int j = i & (half - 1);
            u64 a = x[i], b = x[i + half];
            x[i] = gl_add(a, b);
            u64 d = gl_sub(a, b);
            x[i + half] = (j == 0) ? d : gl_mul(d, tw[(j << (Q - 1 - s)) * tstride]);
        }
    }
}

template <int LOG_R1, int LOG_R2, int LOG_T>
struct PassCfg {
    static constexpr int R1 = 1 << LOG_R1, R2 = 1 << LOG_R2, LOG_R = LOG_R1 + LOG_R2, R = 1 << LOG_R, T = 1 << LOG_T;
    static constexpr int NT = R1 * T;
    static constexpr int PADROW = T > 1 ? 1 : 0;
    static constexpr int ROWW = T + PADROW;
    static constexpr int DATA_ELEMS = R * ROWW + R2 * T + 8;
    static constexpr size_t SMEM = sizeof(u64) * (size_t)(R + DATA_ELEMS) + 16;
    __device__ static __forceinline__ int idx(int row, int col) { return row * ROWW + (row >> LOG_R1) * T + col; }
};

template <int LOG_R1, int LOG_R2, int LOG_T>
__global__ void __launch_bounds__(PassCfg<LOG_R1, LOG_R2, LOG_T>::NT)
    ntt_pass_kernel(const __grid_constant__ PassParams P) {
    using C = PassCfg<LOG_R1, LOG_R2, LOG_T>;
    constexpr int R1 = C::R1, R2 = C::R2, R = C::R, T = C::T, LOG_R = C::LOG_R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *tw_s = reinterpret_cast<u64 *>(smem_raw);  // R entries
    u64 *data = tw_s + R;
    u64 *mbar = data + C::DATA_ELEMS;

    const int tid = threadIdx.x;
    const u32 log_ncols = P.log_L - LOG_R + P.log_I;  // columns per plane = (L/R)*I
    const u64 col0 = (u64)blockIdx.x << LOG_T;
    const u64 *in = P.in + (u64)blockIdx.y * P.in_stride;
    u64 *out = P.out + (u64)blockIdx.y * P.out_stride;

    // ---- stage the twiddle table with one TMA bulk copy ------------------------------
    const u32 mbar_a = smem_u32(mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        constexpr u32 bytes = R * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(tw_s)),
                     "l"(P.tw), "r"(bytes), "r"(mbar_a)
                     : "memory");
    }

    // ---- phase 1: global -> registers, R1-point NTT, twiddle, -> shared ---------------
    const bool p1_active = tid < R2 * T;
    u64 x[R1];
    int jp = 0, col = 0;
    if (p1_active) {
        col = tid & (T - 1);
        jp = tid >> LOG_T;
        const u64 colg = col0 + col;
#pragma unroll
        for (int t = 0; t < R1; ++t) {
            const u64 j = ((u64)(jp + R2 * t) << log_ncols) + colg;
            x[t] = ((P.flags & F_FIRST) && j >= P.n_in) ? 0 : in[j];
        }
        if (P.flags & F_FIRST_SCALE) {
            // offset^j, j = (jp + R2*t)*ncols + colg: geometric in t with ratio offset^(R2*ncols)
            u64 f = gl_pow_sq(P.s_sq, ((u64)jp << log_ncols) + colg);
            const u64 ratio = P.s_sq[LOG_R2 + log_ncols];
#pragma unroll
            for (int t = 0; t < R1; ++t) {
                x[t] = gl_mul(x[t], f);
                f = gl_mul(f, ratio);
            }
        }
    }
    __syncthreads();  // mbarrier init visible to all threads
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_TW:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_TW;\n"
        "bra WAIT_TW;\n"
        "DONE_TW:\n"
        "}\n" ::"r"(mbar_a)
        : "memory");
    if (p1_active) {
        reg_ntt<LOG_R1>(x, tw_s, R2);  // root omega_R^(R/R1)
#pragma unroll
        for (int i = 0; i < R1; ++i) {
            const int k = brev<LOG_R1>(i);
            u64 v = x[i];
            if (LOG_R2 > 0 && k != 0) v = gl_mul(v, tw_s[jp * k]);
            data[C::idx(jp * R1 + k, col)] = v;
        }
    }
    __syncthreads();

    // ---- phase 2: shared -> registers, R2-point NTT, twiddle / scale, -> global -------
    int c;
    if (P.flags & F_CFAST) {
        c = tid & (R1 - 1);
        col = tid >> LOG_R1;
    } else {
        col = tid & (T - 1);
        c = tid >> LOG_T;
    }
    u64 y[R2];
#pragma unroll
    for (int t = 0; t < R2; ++t) y[t] = data[C::idx(c + R1 * t, col)];
    reg_ntt<LOG_R2>(y, tw_s, R1);  // root omega_R^(R/R2)

    const u64 colg = col0 + col;
    const u64 jprime = colg >> P.log_I;
    const u64 cI = colg & (((u64)1 << P.log_I) - 1);
    if (P.log_L > (u32)LOG_R) {
        // inter-pass twiddle omega_L^(j' * k), k = k2*R1 + c: geometric in k2
        const u32 sh = P.log_n - P.log_L;
        u64 wb = gl_pow_sq(P.w_sq, (jprime * (u64)c) << sh);
        const u64 wr = gl_pow_sq(P.w_sq, (jprime << LOG_R1) << sh);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            y[brev<LOG_R2>(k2)] = gl_mul(y[brev<LOG_R2>(k2)], wb);
            wb = gl_mul(wb, wr);
        }
    }
    if (P.flags & F_LAST_SCALE) {
        // out index = (k2*R1 + c)*I + colg; multiplier out_mul * s^index
        u64 f = gl_mul(P.out_mul, gl_pow_sq(P.s_sq, ((u64)c << P.log_I) + colg));
        const u64 ratio = P.s_sq[LOG_R1 + P.log_I];
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            y[brev<LOG_R2>(k2)] = gl_mul(y[brev<LOG_R2>(k2)], f);
            f = gl_mul(f, ratio);
        }
    }
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
        const u64 kr = (u64)k2 * R1 + c;
        out[(((jprime << LOG_R) + kr) << P.log_I) + cI] = y[brev<LOG_R2>(k2)];
    }
}

__global__ void pow_table_kernel(u64 *tab, u64 base, u32 count) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) tab[i] = gl_pow(base, i);
}

// ---- host side ---------------------------------------------------------------------
struct TwEntry {
    u64 *ptr;
    cudaEvent_t ready;
};
std::mutex g_tw_mu;
std::map<std::pair<u64, u32>, TwEntry> g_tw;  // (omega_R, log_R) -> table

int get_tw(u64 omega_R, u32 log_R, cudaStream_t st, const u64 **out) {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    auto key = std::make_pair(omega_R ^ ((u64)dev << 56), log_R);  // per-device tables
    auto it = g_tw.find(key);
    if (it == g_tw.end()) {
        TwEntry e;
        const u32 R = 1u << log_R;
        B2S_CUDA(cudaMalloc(&e.ptr, sizeof(u64) * (R < 2 ? 2 : R)));
        B2S_CUDA(cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
        pow_table_kernel<<<(R + 255) / 256, 256, 0, st>>>(e.ptr, omega_R, R);
        B2S_LAUNCHED();
        B2S_CUDA(cudaEventRecord(e.ready, st));
        it = g_tw.emplace(key, e).first;
    } else {
        B2S_CUDA(cudaStreamWaitEvent(st, it->second.ready, 0));
    }
    *out = it->second.ptr;
    return 0;
}

template <int A, int B, int LT>
int launch_pass(const PassParams &P, u32 tiles, u32 planes, cudaStream_t st) {
    using C = PassCfg<A, B, LT>;
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        B2S_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<A, B, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C::SMEM));
        attr_done[dev & 63] = true;
    }
    ntt_pass_kernel<A, B, LT><<<dim3(tiles, planes), C::NT, C::SMEM, st>>>(P);
    B2S_LAUNCHED();
    return 0;
}

int dispatch_pass(u32 log_R, u32 log_T, const PassParams &P, u32 tiles, u32 planes, cudaStream_t st) {
#define CASE(r, a, b)                                                        \
    case r:                                                                  \
        return log_T == 3 ? launch_pass<a, b, 3>(P, tiles, planes, st)       \
                          : launch_pass<a, b, 0>(P, tiles, planes, st);
    switch (log_R) {
        CASE(1, 1, 0)
        CASE(2, 1, 1)
        CASE(3, 2, 1)
        CASE(4, 2, 2)
        CASE(5, 3, 2)
        CASE(6, 3, 3)
        CASE(7, 4, 3)
        CASE(8, 4, 4)
        CASE(9, 5, 4)
        CASE(10, 5, 5)
    }
#undef CASE
    b2s_set_error("unsupported pass radix 2^%u", log_R);
    return B2S_ERR_ARG;
}

}  // namespace

void ntt_cache_clear() {
    std::lock_guard<std::mutex> lk(g_tw_mu);
    for (auto &kv : g_tw) {
        cudaFree(kv.second.ptr);
        cudaEventDestroy(kv.second.ready);
    }
    g_tw.clear();
}

int ntt_run(const u64 *d_in, u64 in_stride, u32 n_in, u64 *d_out, u64 out_stride, u32 log_n, u32 n_planes, u64 omega,
            u64 offset, int inverse, cudaStream_t st) {
    if (log_n > 30) {
        b2s_set_error("log_n %u too large", log_n);
        return B2S_ERR_ARG;
    }
    const u64 n = (u64)1 << log_n;
    if (n_in > n || (inverse && n_in != n)) {
        b2s_set_error("n_in %u inconsistent with n %llu", n_in, (unsigned long long)n);
        return B2S_ERR_ARG;
    }
    if (n_planes == 0) return 0;
    // code/ntt.py:13-16 (and :29-36 for intt)
    if (gl_pow(omega, n) != 1) {
        b2s_set_error("primitive root must be nth root of unity, where n is %llu", (unsigned long long)n);
        return B2S_ERR_ASSERT_ROOT;
    }
    if (log_n >= 1 && gl_pow(omega, n / 2) == 1) {
        b2s_set_error("primitive root is not primitive nth root of unity, where n is %llu", (unsigned long long)n);
        return B2S_ERR_ASSERT_PRIMITIVE;
    }
    if (log_n == 0) {  // code/ntt.py:8-9 / :32-33: identity
        if (d_in != d_out)
            for (u32 q = 0; q < n_planes; ++q)
                B2S_CUDA(cudaMemcpyAsync(d_out + q * out_stride, d_in + q * in_stride, sizeof(u64),
                                         cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const u64 w = inverse ? gl_inv(omega) : omega;            // code/ntt.py:41
    const u64 scale = inverse ? gl_inv(offset) : offset;      // code/ntt.py:165 / :174
    const bool do_scale = offset != 1;

    // pass plan: digits of log_n, most significant input digit first
    std::vector<u32> radix;
    if (log_n <= 10) {
        radix = {log_n};
    } else if (log_n <= 20) {
        radix = {(log_n + 1) / 2, log_n / 2};
    } else {
        u32 a = (log_n + 2) / 3, b = (log_n - a + 1) / 2;
        radix = {a, b, log_n - a - b};
    }
    const size_t npass = radix.size();

    PassParams P;
    P.in_stride = in_stride;
    P.out_stride = out_stride;
    P.log_n = log_n;
    P.n_in = n_in;
    P.out_mul = inverse ? gl_inv(n % GL_P) : 1;  // code/ntt.py:39
    u64 sq = w, ss = scale;
    for (int b = 0; b < 32; ++b) {
        P.w_sq[b] = sq;
        P.s_sq[b] = ss;
        sq = gl_mul(sq, sq);
        ss = gl_mul(ss, ss);
    }

    // scratch for the intermediate vectors (stream-ordered pool)
    u64 *work[2] = {nullptr, nullptr};
    for (size_t i = 0; i + 1 < npass; ++i)
        B2S_CUDA(cudaMallocAsync(&work[i], sizeof(u64) * n * n_planes, st));

    u32 log_L = log_n, log_I = 0;
    int rc = 0;
    for (size_t ps = 0; ps < npass && rc == 0; ++ps) {
        const u32 log_R = radix[ps];
        const bool first = ps == 0, last = ps + 1 == npass;
        P.in = first ? d_in : work[ps - 1];
        P.in_stride = first ? in_stride : n;
        P.out = last ? d_out : work[ps];
        P.out_stride = last ? out_stride : n;
        P.log_L = log_L;
        P.log_I = log_I;
        const u32 log_ncols = log_L - log_R + log_I;
        const u32 log_T = (npass > 1 && log_ncols >= 3) ? 3 : 0;
        P.flags = 0;
        if (first) P.flags |= F_FIRST;
        if (first && do_scale && !inverse) P.flags |= F_FIRST_SCALE;
        if (last && inverse) P.flags |= F_LAST_SCALE;  // n^-1 (and offset^-k when offset != 1)
        if (log_I == 0 && log_T != 0) P.flags |= F_CFAST;
        rc = get_tw(gl_pow(w, n >> log_R), log_R, st, &P.tw);
        if (rc) break;
        rc = dispatch_pass(log_R, log_T, P, (u32)(((u64)1 << log_ncols) >> log_T), n_planes, st);
        log_L -= log_R;
        log_I += log_R;
    }
    for (size_t i = 0; i + 1 < npass; ++i)
        if (work[i]) cudaFreeAsync(work[i], st);
    return rc;
}
