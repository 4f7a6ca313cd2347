// ntt.cu -- batched number-theoretic transforms over F_p (p = 2^64 - 2^32 + 1).
//
// Replaces code/ntt.py:4-42 (ntt/intt) and the scale+pad+ntt wrappers of code/ntt.py:164-174
// and code/fri.py:26-44.  Natural order in, natural order out:
//     out[k] = sum_j in[j] * omega^(j*k)
//
// Schedule (not the reference's recursion).  A transform of length n = 2^log_n is split into
// 1-3 PASSES over global memory (Cooley-Tukey on the index digits, most significant input
// digit first).  One CTA owns a TILE of R rows x T columns (R <= 1024 = the radix of the
// pass, T adjacent columns so that every global access is a >= 32-byte segment), keeps it in
// shared memory and runs all log2(R) butterfly stages there:
//
//   load   global -> shared, fused with the zero padding beyond n_in and the row part of the
//          coset scale offset^j (first pass); non-last passes load rows in bit-reversed order
//          (rows of a tile are separate global segments, so permuting them is free)
//   steps  radix-8/16 register butterflies over shared memory (in place, 8-16 elements per
//          thread per step, twiddles from a table staged by one TMA bulk copy):
//          decimation in time for non-last passes (natural-order result), decimation in
//          frequency for the last pass (bit-reversed rows, undone by the row permutation of
//          the store)
//   store  shared -> global, fused with the inter-pass twiddle omega^(col*k) (two-level table
//          built per CTA), the column part of the coset scale, and for the inverse transform
//          n^-1 * offset^-k; only the last pass canonicalises.
//
// The intermediate vector keeps the tile shape (pass p writes where it read), so one scratch
// buffer suffices and, with the data resident in the 126 MB L2, HBM sees only the compulsory
// read of the input and write of the output.
#include <map>
#include <mutex>
#include <vector>

#include "common.h"
#include "glfast.cuh"

namespace {

enum : u32 {
    F_FIRST = 1,         // bounds check against n_in (zero padding)
    F_TWIDDLE = 2,       // non-last pass: multiply by omega^(tw_mul * col * k)
    F_COLSCALE = 4,      // fold scale^col into the inter-pass twiddle (forward coset, first pass)
    F_OUT_MUL = 8,       // multiply the output by out_mul (n^-1)
    F_LOAD_ROWFAST = 16  // rows are the contiguous global dimension of the input tile
};

struct PassParams {
    const u64 *in;
    u64 *out;
    u64 in_plane_stride, out_plane_stride;
    u64 in_blk_stride, out_blk_stride;
    u64 in_row_stride, in_col_stride, out_row_stride;  // output columns are always contiguous
    const u64 *tw;         // omega_R^e, e < R
    const u64 *in_scale;   // (scale^in_row_stride)^r, r < R, or null
    const u64 *out_scale;  // (scale^out_row_stride)^k, k < R, or null
    u64 out_mul;
    u64 tw_mul;
    u64 n_in;
    u32 flags;
    u32 log_ncols;
    u64 w_sq[32];  // omega^(2^b)
    u64 s_sq[32];  // scale^(2^b)
};

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

template <int LOG_R, int LOG_T>
struct Cfg {
    static constexpr int R = 1 << LOG_R, T = 1 << LOG_T;
    static constexpr int NSTEPS = (LOG_R + 3) / 4;  // steps of radix <= 16
    static constexpr int QMAX = (LOG_R + NSTEPS - 1) / NSTEPS;
    static constexpr int NT_RAW = (R * T) >> (QMAX < 3 ? QMAX : 3);
    static constexpr int NT = NT_RAW < 32 ? 32 : (NT_RAW > 1024 ? 1024 : NT_RAW);
    static constexpr int PS = LOG_T == 0 ? 4 : 2;  // row padding shift (bank-conflict search)
    static constexpr int DATA = R * T + (R >> PS) + 8;
    static constexpr int UV = 64 * T;
    static constexpr size_t SMEM = sizeof(u64) * (size_t)(R + DATA + UV) + 16;
    __device__ static __forceinline__ int idx(int row, int col) { return (row << LOG_T) + col + (row >> PS); }
    // step s handles q(s) stages; the first steps take the larger radices
    __host__ __device__ static constexpr int q_of(int s) { return LOG_R / NSTEPS + (s < LOG_R % NSTEPS ? 1 : 0); }
};

// One radix-2^Q step on the shared-memory tile.  The 2^Q elements of an item sit at rows
// i0 + k*hq (k < 2^Q); LOG_HQ = log2(hq) is the lowest butterfly half-distance of the step.
template <int LOG_R, int LOG_T, int Q, int LOG_HQ, bool DIT>
__device__ __forceinline__ void radix_step(u64 *data, const u64 *tw, int tid) {
    using C = Cfg<LOG_R, LOG_T>;
    constexpr int R = C::R, T = C::T, NQ = 1 << Q, HQ = 1 << LOG_HQ;
    constexpr int ITEMS = (R * T) >> Q;
    for (int u = tid; u < ITEMS; u += C::NT) {
        const int col = u & (T - 1);
        const int g = u >> LOG_T;
        const int r = g & (HQ - 1);
        const int i0 = ((g >> LOG_HQ) << (LOG_HQ + Q)) | r;
        u64 x[NQ];
#pragma unroll
        for (int k = 0; k < NQ; ++k) x[k] = data[C::idx(i0 + (k << LOG_HQ), col)];
        if (DIT) {
            // stage t: half H = hq*2^t, pairs (k, k + 2^t), twiddle omega_R^((r + kk*hq) * R/(2H))
#pragma unroll
            for (int t = 0; t < Q; ++t) {
                const int log_m = LOG_R - 1 - LOG_HQ - t;  // log2(R / (2H))
#pragma unroll
                for (int k = 0; k < NQ; ++k) {
                    if (k & (1 << t)) continue;
                    const int kk = k & ((1 << t) - 1);
                    u64 b = x[k + (1 << t)];
                    if (LOG_HQ + t > 0 || kk > 0) {  // H == 1: the only twiddle is 1
                        const int e = (r << log_m) + (kk << (LOG_R - 1 - t));
                        b = fmul(b, tw[e]);
                    }
                    b = canon(b);
                    const u64 a = x[k];
                    x[k] = fadd(a, b);
                    x[k + (1 << t)] = fsub(a, b);
                }
            }
        } else {
            // stage t: half H = hq*2^(Q-1-t), pairs (k, k + 2^(Q-1-t))
#pragma unroll
            for (int t = 0; t < Q; ++t) {
                const int sh = Q - 1 - t;
                const int log_m = LOG_R - 1 - LOG_HQ - sh;  // log2(R / (2H))
#pragma unroll
                for (int k = 0; k < NQ; ++k) {
                    if (k & (1 << sh)) continue;
                    const int kk = k & ((1 << sh) - 1);
                    const u64 a = x[k];
                    const u64 b = canon(x[k + (1 << sh)]);
                    x[k] = fadd(a, b);
                    u64 d = fsub(a, b);
                    if (LOG_HQ + sh > 0 || kk > 0) {
                        const int e = (r << log_m) + (kk << (LOG_R - 1 - sh));
                        d = fmul(d, tw[e]);
                    }
                    x[k + (1 << sh)] = d;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NQ; ++k) data[C::idx(i0 + (k << LOG_HQ), col)] = x[k];
    }
}

// all steps of the tile transform, unrolled at compile time
template <int LOG_R, int LOG_T, bool DIT, int S, int DONE>
__device__ __forceinline__ void run_steps(u64 *data, const u64 *tw, int tid) {
    using C = Cfg<LOG_R, LOG_T>;
    if constexpr (S < C::NSTEPS) {
        constexpr int Q = C::q_of(S);
        // DIT walks the half-distances upwards (hq = 2^DONE), DIF downwards
        constexpr int LOG_HQ = DIT ? DONE : LOG_R - DONE - Q;
        radix_step<LOG_R, LOG_T, Q, LOG_HQ, DIT>(data, tw, tid);
        if (S + 1 < C::NSTEPS) __syncthreads();
        run_steps<LOG_R, LOG_T, DIT, S + 1, DONE + Q>(data, tw, tid);
    }
}

template <int LOG_R, int LOG_T, bool DIT>
__global__ void __launch_bounds__(Cfg<LOG_R, LOG_T>::NT) ntt_pass_kernel(const __grid_constant__ PassParams P) {
    using C = Cfg<LOG_R, LOG_T>;
    constexpr int R = C::R, T = C::T, NT = C::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *tw_s = reinterpret_cast<u64 *>(smem_raw);  // R entries (TMA destination, 16-byte aligned)
    u64 *data = tw_s + R;
    u64 *U = data + C::DATA;  // [T][32]  omega^(tw_mul*col*i) (* scale^col)
    u64 *V = U + 32 * T;      // [T][32]  omega^(tw_mul*col*32*i)
    u64 *mbar = V + 32 * T;

    const int tid = threadIdx.x;
    const u64 col0 = (u64)blockIdx.x << LOG_T;
    const u64 blk = blockIdx.y;
    const u64 *in = P.in + (u64)blockIdx.z * P.in_plane_stride + blk * P.in_blk_stride;
    u64 *out = P.out + (u64)blockIdx.z * P.out_plane_stride + blk * P.out_blk_stride;

    // ---- twiddle table: one TMA bulk copy, completion on an mbarrier -------------------
    const u32 mbar_a = smem_u32(mbar);
    if (tid == 0) {
        constexpr u32 bytes = (R < 2 ? 2 : R) * 8;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(tw_s)),
                     "l"(P.tw), "r"(bytes), "r"(mbar_a)
                     : "memory");
    }

    // ---- load the tile -------------------------------------------------------------------
    const bool rowfast = (P.flags & F_LOAD_ROWFAST) != 0;
    for (int e = tid; e < R * T; e += NT) {
        int row, col;
        if (rowfast) {
            row = e & (R - 1);
            col = e >> LOG_R;
        } else {
            col = e & (T - 1);
            row = e >> LOG_T;
        }
        const int src = DIT ? (int)(__brev((unsigned)row) >> (32 - LOG_R)) : row;
        const u64 j = (u64)src * P.in_row_stride + (col0 + col) * P.in_col_stride;
        u64 v = 0;
        if (!(P.flags & F_FIRST) || j < P.n_in) v = in[j];
        if (P.in_scale) v = fmul(v, P.in_scale[src]);
        data[C::idx(row, col)] = v;
    }

    // ---- per-column inter-pass twiddle tables (two-level: k = 32*hi + lo) -------------------
    if (P.flags & F_TWIDDLE) {
        for (int t = tid; t < 64 * T; t += NT) {
            const int c = t >> 6, i = t & 63;
            const u64 colg = col0 + c;
            u64 v;
            if (i < 32) {
                v = fpow_sq(P.w_sq, P.tw_mul * colg * (u64)i);
                if (P.flags & F_COLSCALE) v = fmul(v, fpow_sq(P.s_sq, colg));
                U[c * 32 + i] = v;
            } else {
                V[c * 32 + (i - 32)] = fpow_sq(P.w_sq, P.tw_mul * colg * (u64)(32 * (i - 32)));
            }
        }
    }
    __syncthreads();  // tile + tables written, mbarrier initialised
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

    // ---- butterflies in shared memory -----------------------------------------------------------
    run_steps<LOG_R, LOG_T, DIT, 0, 0>(data, tw_s, tid);
    __syncthreads();

    // ---- store -----------------------------------------------------------------------------
    // NT is a multiple of T, so a thread keeps its column through the loop
    const int col = tid & (T - 1);
    const u64 colg = col0 + col;
    u64 cconst = 1;
    if (P.out_scale) cconst = fmul(P.out_mul, fpow_sq(P.s_sq, blk * P.out_blk_stride + colg));
    for (int e = tid; e < R * T; e += NT) {
        const int row = e >> LOG_T;
        const int k = DIT ? row : (int)(__brev((unsigned)row) >> (32 - LOG_R));
        u64 v = data[C::idx(row, col)];
        if (P.flags & F_TWIDDLE) v = fmul(v, fmul(U[col * 32 + (k & 31)], V[col * 32 + (k >> 5)]));
        if (P.out_scale)
            v = fmul(fmul(v, cconst), P.out_scale[k]);
        else if (P.flags & F_OUT_MUL)
            v = fmul(v, P.out_mul);
        if (!DIT) v = canon(v);  // last pass: canonical integers leave the library
        out[(u64)k * P.out_row_stride + colg] = v;
    }
}

__global__ void pow_table_kernel(u64 *tab, u64 base, u32 count) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) tab[i] = gl_pow(base, i);
}

// ---- host side -------------------------------------------------------------------------
struct TabEntry {
    u64 *ptr;
    cudaEvent_t ready;
};
std::mutex g_tab_mu;
std::map<std::pair<u64, u32>, TabEntry> g_tab;  // (base, log_count) -> base^i, i < 2^log_count

int get_pow_table(u64 base, u32 log_count, cudaStream_t st, const u64 **out) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    auto key = std::make_pair(base, log_count | ((u32)dev << 8));
    auto it = g_tab.find(key);
    if (it == g_tab.end()) {
        TabEntry e;
        const u32 cnt = 1u << log_count;
        B2S_CUDA(cudaMalloc(&e.ptr, sizeof(u64) * (cnt < 2 ? 2 : cnt)));
        B2S_CUDA(cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
        pow_table_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(e.ptr, base, cnt < 2 ? 2 : cnt);
        B2S_LAUNCHED();
        B2S_CUDA(cudaEventRecord(e.ready, st));
        it = g_tab.emplace(key, e).first;
    } else {
        B2S_CUDA(cudaStreamWaitEvent(st, it->second.ready, 0));
    }
    *out = it->second.ptr;
    return 0;
}

template <int LR, int LT, bool DIT>
int launch_pass(const PassParams &P, dim3 grid, cudaStream_t st) {
    using C = Cfg<LR, LT>;
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 15]) {
        B2S_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<LR, LT, DIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C::SMEM));
        attr_done[dev & 15] = true;
    }
    ntt_pass_kernel<LR, LT, DIT><<<grid, C::NT, C::SMEM, st>>>(P);
    B2S_LAUNCHED();
    return 0;
}

constexpr int TILE_LOG_T = 2;  // 4 columns = 32-byte segments

int dispatch_pass(u32 log_R, u32 log_T, bool dit, const PassParams &P, dim3 grid, cudaStream_t st) {
#define CASE(r)                                                                                       \
    case r:                                                                                           \
        if (log_T == 0) return dit ? launch_pass<r, 0, true>(P, grid, st) : launch_pass<r, 0, false>(P, grid, st); \
        return dit ? launch_pass<r, TILE_LOG_T, true>(P, grid, st) : launch_pass<r, TILE_LOG_T, false>(P, grid, st);
    switch (log_R) {
        CASE(1)
        CASE(2)
        CASE(3)
        CASE(4)
        CASE(5)
        CASE(6)
        CASE(7)
        CASE(8)
        CASE(9)
        CASE(10)
    }
#undef CASE
    b2s_set_error("unsupported pass radix 2^%u", log_R);
    return B2S_ERR_ARG;
}

}  // namespace

void ntt_cache_clear() {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    for (auto &kv : g_tab) {
        cudaFree(kv.second.ptr);
        cudaEventDestroy(kv.second.ready);
    }
    g_tab.clear();
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
    if (n_planes > 65535) {
        b2s_set_error("at most 65535 planes per call");
        return B2S_ERR_ARG;
    }
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
    const u64 w = inverse ? gl_inv(omega) : omega;        // code/ntt.py:41
    const u64 scale = inverse ? gl_inv(offset) : offset;  // code/ntt.py:165 / :174
    const bool do_scale = offset != 1;

    // digits of log_n: lg[0] is the radix of the first pass (most significant input digit)
    u32 lg[3] = {0, 0, 0};
    int npass;
    if (log_n <= 10) {
        npass = 1;
        lg[0] = log_n;
    } else if (log_n <= 20) {
        npass = 2;
        lg[0] = (log_n + 1) / 2;
        lg[1] = log_n / 2;
    } else {
        npass = 3;
        lg[0] = (log_n + 2) / 3;
        lg[1] = (log_n - lg[0] + 1) / 2;
        lg[2] = log_n - lg[0] - lg[1];
    }

    PassParams P;
    P.n_in = n_in;
    u64 sq = w, ss = scale;
    for (int b = 0; b < 32; ++b) {
        P.w_sq[b] = sq;
        P.s_sq[b] = ss;
        sq = gl_mul(sq, sq);
        ss = gl_mul(ss, ss);
    }
    const u64 ninv = gl_inv(n % GL_P);  // code/ntt.py:39

    u64 *work = nullptr;
    if (npass > 1) B2S_CUDA(cudaMallocAsync(&work, sizeof(u64) * n * n_planes, st));

    int rc = 0;
    u32 rest = log_n;  // log2 of the product of the radices not yet processed, including this pass
    for (int ps = 0; ps < npass && rc == 0; ++ps) {
        const u32 log_R = lg[ps];
        rest -= log_R;  // log2(columns of this pass's tiles within a block)
        const bool first = ps == 0, last = ps + 1 == npass;
        const u32 log_T = (npass == 1) ? 0 : TILE_LOG_T;
        P.in = first ? d_in : work;
        P.in_plane_stride = first ? in_stride : n;
        P.out = last ? d_out : work;
        P.out_plane_stride = last ? out_stride : n;
        P.flags = first ? F_FIRST : 0;
        P.in_scale = P.out_scale = nullptr;
        P.out_mul = 1;
        P.tw_mul = 1;
        dim3 grid(1, 1, n_planes);
        if (!last) {
            // rows = current top digit (stride 2^rest), columns = all lower digits, same shape out
            // pass 0 works on the whole vector, pass 1 of a 3-pass plan on each block k3
            const u64 ncols = (u64)1 << rest;
            P.in_row_stride = P.out_row_stride = ncols;
            P.in_col_stride = 1;
            P.in_blk_stride = P.out_blk_stride = ps == 0 ? 0 : ((u64)1 << (log_n - lg[0]));
            P.log_ncols = rest;
            P.flags |= F_TWIDDLE;
            P.tw_mul = ps == 0 ? 1 : ((u64)1 << lg[0]);  // omega_L = omega^(n/L)
            if (first && do_scale && !inverse) {
                P.flags |= F_COLSCALE;
                rc = get_pow_table(gl_pow(scale, ncols), log_R, st, &P.in_scale);
                if (rc) break;
            }
            grid.x = (unsigned)(ncols >> log_T);
            grid.y = ps == 0 ? 1 : (1u << lg[0]);
        } else if (npass == 1) {
            P.in_row_stride = P.out_row_stride = 1;
            P.in_col_stride = 0;
            P.in_blk_stride = P.out_blk_stride = 0;
            P.log_ncols = 0;
            if (do_scale && !inverse) {
                rc = get_pow_table(scale, log_R, st, &P.in_scale);
                if (rc) break;
            }
        } else {
            // last pass of a multi-pass plan: rows = lowest input digit j1 (contiguous), tile
            // columns = T adjacent values of the FIRST pass's output digit, blocks = middle digit
            const u32 log_n1 = lg[npass - 1];
            P.flags |= F_LOAD_ROWFAST;
            P.in_row_stride = 1;
            if (npass == 2) {
                P.in_col_stride = (u64)1 << log_n1;  // k2 * n1
                P.in_blk_stride = P.out_blk_stride = 0;
                P.out_row_stride = (u64)1 << lg[0];  // k1 * n2
                grid.x = (1u << lg[0]) >> log_T;
            } else {
                P.in_col_stride = (u64)1 << (lg[1] + lg[2]);  // k3 * n1*n2
                P.in_blk_stride = (u64)1 << lg[2];            // k2 * n1
                P.out_blk_stride = (u64)1 << lg[0];           // k2 * n3
                P.out_row_stride = (u64)1 << (lg[0] + lg[1]);  // k1 * n2*n3
                grid.x = (1u << lg[0]) >> log_T;
                grid.y = 1u << lg[1];
            }
            P.log_ncols = lg[0];
        }
        if (last && inverse) {
            P.out_mul = ninv;
            P.flags |= F_OUT_MUL;
            if (do_scale) {
                rc = get_pow_table(gl_pow(scale, P.out_row_stride), log_R, st, &P.out_scale);
                if (rc) break;
            }
        }
        rc = get_pow_table(gl_pow(w, n >> log_R), log_R, st, &P.tw);
        if (rc) break;
        rc = dispatch_pass(log_R, log_T, !last, P, grid, st);
    }
    if (work) cudaFreeAsync(work, st);
    return rc;
}
