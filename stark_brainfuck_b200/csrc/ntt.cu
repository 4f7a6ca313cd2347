// ntt.cu -- batched number-theoretic transforms over F_p (p = 2^64 - 2^32 + 1).
//
// Replaces code/ntt.py:4-42 (ntt/intt) and the scale+pad+ntt wrappers of code/ntt.py:164-174
// and code/fri.py:26-44.  Natural order in, natural order out:
//     out[k] = sum_j in[j] * omega^(j*k)
//
// Schedule (not the reference's recursion): a transform of length n = 2^log_n is split into
// 1-3 PASSES over global memory (Cooley-Tukey on the index digits, most significant input
// digit first; plan: ntt4_plan.h).  One CTA owns a tile of R rows x T columns (R = the radix of
// the pass, up to 2048; T = 4 adjacent columns so that every global access is a full 32-byte
// sector), keeps it in shared memory and runs the three compact phases of ntt4.cuh on it.  The
// intermediate vector keeps the tile shape (a non-last pass writes where it read), so one
// scratch buffer suffices and, with the data resident in the 126 MB L2, HBM sees only the
// compulsory read of the input and write of the output.  Twiddle tables are built once per
// (root, size) on the device, cached for the life of the library and staged into shared memory
// by TMA bulk copies.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "common.h"
#include "ntt4_plan.h"

namespace {

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// WIDE = true: the same pass compiled for at most 256 threads per CTA and up to 128 registers per thread.  A lone
// vector puts only ~14 warps on an SM, so registers are free and the scheduler may keep more butterflies in flight.
template <int TL, int LE, bool WIDE>
__global__ void __launch_bounds__(WIDE ? 256 : (LE == 4 ? 512 : 1024), WIDE ? 2 : (LE == 4 ? 2 : 1))
    ntt4_pass_kernel(const __grid_constant__ Pass4Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 R = 1u << P.log_R;
    const u32 n_core0 = pass4_core_table_elems(LE, P.a, 0), n_core1 = pass4_core_table_elems(LE, P.a, 1);
    u64 *tw_tail_s = reinterpret_cast<u64 *>(smem_raw);  // R entries when TL > 0          (TMA destinations,
    u64 *tw_core_s = tw_tail_s + (TL > 0 ? R : 0);       // tables of core steps 0 .. a-2   16-byte aligned)
    u64 *S = tw_core_s + n_core0 + n_core1;
    u64 *mbar = S + ((size_t)P.cs << P.log_T);
    const u32 tid = threadIdx.x, nthreads = blockDim.x;

    // twiddle tables of this pass: TMA bulk copies, completion on one mbarrier
    const u32 mbar_a = smem_u32(mbar);
    const u32 bytes_tail = TL > 0 ? R * 8 : 0, bytes0 = n_core0 * 8, bytes1 = n_core1 * 8;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes_tail + bytes0 + bytes1)
                     : "memory");
        if (bytes_tail)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tw_tail_s)),
                         "l"(P.tw_tail), "r"(bytes_tail), "r"(mbar_a)
                         : "memory");
        if (bytes0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tw_core_s)),
                         "l"(P.tw_core[0]), "r"(bytes0), "r"(mbar_a)
                         : "memory");
        if (bytes1)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tw_core_s + n_core0)),
                         "l"(P.tw_core[1]), "r"(bytes1), "r"(mbar_a)
                         : "memory");
    }
    __syncthreads();  // mbarrier initialised
    // Programmatic dependent launch: a non-first pass is launched while its predecessor drains;
    // everything above (and the launch latency) overlaps it, the data it reads does not.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    auto tables_ready = [mbar_a] {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_TW4:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra DONE_TW4;\n"
            "bra WAIT_TW4;\n"
            "DONE_TW4:\n"
            "}\n" ::"r"(mbar_a)
            : "memory");
    };
    pass4_tail<TL, LE>(P, tid, nthreads, blockIdx.x, blockIdx.y, blockIdx.z, tw_tail_s, S, tables_ready);
    __syncthreads();
    for (u32 s = 0; s < P.a; ++s) {
        pass4_core<LE>(P, s, tid, nthreads, tw_core_s, S);
        if (s + 1 < P.a) __syncthreads();
    }
    asm volatile("griddepcontrol.launch_dependents;");
    // the out phase reads back exactly what this thread wrote in the last core step
    pass4_out<LE>(P, tid, nthreads, blockIdx.x, blockIdx.y, blockIdx.z, S);
}

// tab[i] = Montgomery form of base^i (kind 1) or of base^((i >> log_l) * (i mod 2^log_l)) (kind 2:
// the [k][lo] twiddle table of a step)
__global__ void table4_kernel(u64 *tab, u64 base, u32 count, u32 kind, u32 log_l) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u64 e = kind == 2 ? (u64)(i >> log_l) * (i & ((1u << log_l) - 1)) : i;
    tab[i] = gl_to_mont(gl_pow(base, e));
}

// lengths 2, 4, 8: one thread per output, out[k] = sum_j scale^j in[j] w^(jk) (* out_mul * oscale^k)
__global__ void small_dft_kernel(const u64 *in, u64 in_stride, u32 n_in, u64 *out, u64 out_stride, u32 n, u64 w, u64 scale,
                                 u64 out_mul, u64 oscale) {
    const u32 k = threadIdx.x;
    const u64 *x = in + (u64)blockIdx.x * in_stride;
    u64 acc = 0;
    if (k < n) {
        const u64 wk = gl_pow(w, k);
        u64 wjk = 1, sj = 1;
        for (u32 j = 0; j < n_in; ++j) {
            acc = gl_add(acc, gl_mul(gl_mul(x[j], sj), wjk));
            wjk = gl_mul(wjk, wk);
            sj = gl_mul(sj, scale);
        }
        acc = gl_mul(gl_mul(acc, out_mul), gl_pow(oscale, k));
    }
    __syncthreads();  // in == out is allowed
    if (k < n) out[(u64)blockIdx.x * out_stride + k] = acc;
}

// ---- host side -------------------------------------------------------------------------
struct TabEntry {
    u64 *ptr;
};
struct TabKey {
    u64 base;
    u32 log_count, kind, log_l;
    int dev;
    bool operator<(const TabKey &o) const {
        if (base != o.base) return base < o.base;
        if (log_count != o.log_count) return log_count < o.log_count;
        if (kind != o.kind) return kind < o.kind;
        if (log_l != o.log_l) return log_l < o.log_l;
        return dev < o.dev;
    }
};
std::mutex g_tab_mu;
std::map<TabKey, TabEntry> g_tab;  // built once per (base, size, layout, device)
std::atomic<u64> g_tab_generation{0};  // bumped whenever the tables are dropped: cached plans of older generations are stale
size_t g_tab_bytes = 0;            // ... and bounded: callers with ever-changing roots or coset offsets
constexpr size_t TAB_BUDGET_BYTES = (size_t)256 << 20;

int get_table(const Tab4 &t, cudaStream_t st, const u64 **out) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    const TabKey key{t.base, t.log_count, t.two_d ? 2u : 1u, t.two_d ? t.log_r2 : 0u, dev};
    auto it = g_tab.find(key);
    if (it == g_tab.end()) {
        TabEntry e;
        const u32 cnt = 1u << t.log_count;
        const u32 alloc = cnt < 2 ? 2 : cnt;
        if (g_tab_bytes + sizeof(u64) * alloc > TAB_BUDGET_BYTES) {
            // over budget: drop every table (cudaFree waits for the kernels that still read them).  Plans hold
            // table pointers, so they go too; the caller of get_plan rebuilds what it needs.
            for (auto &kv : g_tab) cudaFree(kv.second.ptr);
            g_tab.clear();
            g_tab_bytes = 0;
            g_tab_generation.fetch_add(1);
        }
        B2S_CUDA(cudaMalloc(&e.ptr, sizeof(u64) * alloc));
        g_tab_bytes += sizeof(u64) * alloc;
        table4_kernel<<<(alloc + 255) / 256, 256, 0, st>>>(e.ptr, t.base, alloc, key.kind, key.log_l);
        B2S_LAUNCHED();
        // once per table: block until it is built, so that later calls on ANY stream can use it
        // without a cross-stream dependency (a cudaStreamWaitEvent per table per call costs more
        // host time than a whole 2^20 transform takes on the device)
        B2S_CUDA(cudaStreamSynchronize(st));
        it = g_tab.emplace(key, e).first;
    }
    *out = it->second.ptr;
    return 0;
}

struct PlanKey {
    u64 w, scale, n_in;
    u32 log_n, flags;  // flags: inverse, do_scale, log_E
    int dev;
    bool operator<(const PlanKey &o) const {
        if (w != o.w) return w < o.w;
        if (scale != o.scale) return scale < o.scale;
        if (n_in != o.n_in) return n_in < o.n_in;
        if (log_n != o.log_n) return log_n < o.log_n;
        if (flags != o.flags) return flags < o.flags;
        return dev < o.dev;
    }
};
struct CachedPlan {
    Pass4Plan plan[3];
    int npass;
    u64 generation;  // of the table cache its pointers come from
};
std::mutex g_plan_mu;
std::map<PlanKey, CachedPlan> g_plans;

// 4 columns per tile = full 32-byte sectors (narrower tiles balance a single 2^20 vector better
// over 148 SMs but measured slower: profiles/r01_ntt_experiments.md)
int get_plan(u32 log_n, u64 n_in, u64 w, u64 scale, bool inverse, bool do_scale, u32 log_E, cudaStream_t st,
             Pass4Plan plan[3], int *npass) {
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    const PlanKey key{w, do_scale ? scale : 1, n_in, log_n, (inverse ? 1u : 0u) | (do_scale ? 2u : 0u) | (log_E << 2), dev};
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end() && it->second.generation != g_tab_generation.load()) {
        g_plans.erase(it);  // its tables were dropped (table-cache budget)
        it = g_plans.end();
    }
    if (it == g_plans.end()) {
        if (g_plans.size() > 4096) g_plans.clear();  // bounded: callers with ever-changing roots
        CachedPlan cp;
        cp.npass = plan4(log_n, n_in, w, scale, inverse, do_scale, 2, log_E, cp.plan);
        int rc = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {  // a table that overflows the budget drops the earlier ones
            cp.generation = g_tab_generation.load();
            for (int ps = 0; ps < cp.npass; ++ps) {
                Pass4Plan &pl = cp.plan[ps];
                auto bind = [&](const Tab4 &t, const u64 *&dst) {
                    if (t.used && rc == 0) rc = get_table(t, st, &dst);
                };
                bind(pl.tw_tail, pl.P.tw_tail);
                bind(pl.tw_core[0], pl.P.tw_core[0]);
                bind(pl.tw_core[1], pl.P.tw_core[1]);
                bind(pl.in_scale, pl.P.in_scale);
                bind(pl.out_scale, pl.P.out_scale);
                bind(pl.tw_lo, pl.P.tw_lo);
                bind(pl.tw_hi, pl.P.tw_hi);
                bind(pl.col_scale, pl.P.col_scale);
            }
            if (rc || cp.generation == g_tab_generation.load()) break;
        }
        if (rc) return rc;
        it = g_plans.emplace(key, cp).first;
    }
    for (int ps = 0; ps < 3; ++ps) plan[ps] = it->second.plan[ps];  // copied under the lock
    *npass = it->second.npass;
    return 0;
}

template <int TL, int LE, bool WIDE>
int launch_pass4w(const Pass4Plan &pl, u32 n_planes, cudaStream_t st) {
    const Pass4Params &P = pl.P;
    const size_t smem = sizeof(u64) * ((TL > 0 ? ((size_t)1 << P.log_R) : 0) + pass4_core_table_elems(LE, P.a, 0) +
                                       pass4_core_table_elems(LE, P.a, 1) + pass4_smem_elems(P.log_R, P.log_T, LE)) + 16;
    static size_t attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_done[dev & 15] < smem) {
        B2S_CUDA(cudaFuncSetAttribute(ntt4_pass_kernel<TL, LE, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        attr_done[dev & 15] = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid_x, pl.grid_y, n_planes);
    cfg.blockDim = dim3(pass4_threads(P.log_R, P.log_T, LE));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Every pass, the first one included, is a programmatic dependent launch: what precedes griddepcontrol.wait in
    // the kernel (barrier set-up, TMA of the constant twiddle tables) touches nothing an earlier kernel of the stream
    // produces, and the wait itself returns only when that kernel has completed and flushed.
    static const char *pdl_first = getenv("B2S_NTT_PDL_FIRST");
    cfg.numAttrs = (pl.first && pdl_first && pdl_first[0] == '0') ? 0 : 1;
    B2S_CUDA(cudaLaunchKernelEx(&cfg, ntt4_pass_kernel<TL, LE, WIDE>, P));
    B2S_LAUNCHED();
    return 0;
}

template <int TL, int LE>
int launch_pass4(const Pass4Plan &pl, u32 n_planes, cudaStream_t st) {
    static const char *force = getenv("B2S_NTT_WIDE");
    const u32 threads = pass4_threads(pl.P.log_R, pl.P.log_T, LE);
    const u64 ctas = (u64)pl.grid_x * pl.grid_y * n_planes;
    const bool wide = force ? force[0] == '1' : (threads <= 256 && ctas <= 3 * 148);  // (3 x 2^18: 21.6 vs 22.6 us)
    if (wide && threads <= 256) return launch_pass4w<TL, LE, true>(pl, n_planes, st);
    return launch_pass4w<TL, LE, false>(pl, n_planes, st);
}

int dispatch_pass4(const Pass4Plan &pl, u32 n_planes, cudaStream_t st) {
    if (pl.P.log_E == 4) {
        switch (pl.tail) {
            case 0: return launch_pass4<0, 4>(pl, n_planes, st);
            case 1: return launch_pass4<1, 4>(pl, n_planes, st);
            case 2: return launch_pass4<2, 4>(pl, n_planes, st);
            case 3: return launch_pass4<3, 4>(pl, n_planes, st);
        }
    } else if (pl.P.log_E == 3) {
        switch (pl.tail) {
            case 0: return launch_pass4<0, 3>(pl, n_planes, st);
            case 1: return launch_pass4<1, 3>(pl, n_planes, st);
            case 2: return launch_pass4<2, 3>(pl, n_planes, st);
        }
    }
    b2s_set_error("unsupported pass shape: tail 2^%u, core 2^%u", pl.tail, pl.P.log_E);
    return B2S_ERR_ARG;
}

// The intermediate vector of a multi-pass transform.  One buffer per (device, stream), kept between calls: calls on a
// stream are ordered, so they can share it, and without an allocation / release pair between two transforms the
// first pass of the next one is launched right behind the last pass of this one (programmatic dependent launch).
struct WorkBuf {
    u64 *ptr = nullptr;
    size_t bytes = 0;
};
std::mutex g_work_mu;
std::map<std::pair<int, cudaStream_t>, WorkBuf> g_work;

int get_work(size_t bytes, cudaStream_t st, u64 **out) {
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_work_mu);
    if (g_work.size() > 64) {  // many short-lived streams: do not accumulate their buffers
        for (auto &kv : g_work) cudaFreeAsync(kv.second.ptr, kv.first.second);
        g_work.clear();
    }
    WorkBuf &w = g_work[{dev, st}];
    if (w.bytes < bytes) {
        if (w.ptr) cudaFreeAsync(w.ptr, st);
        w.ptr = nullptr;
        w.bytes = 0;
        B2S_CUDA(cudaMallocAsync(&w.ptr, bytes, st));
        w.bytes = bytes;
    }
    *out = w.ptr;
    return 0;
}

}  // namespace

void ntt_cache_clear() {
    {
        std::lock_guard<std::mutex> lk(g_work_mu);
        for (auto &kv : g_work) cudaFree(kv.second.ptr);
        g_work.clear();
    }
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        g_plans.clear();
    }
    std::lock_guard<std::mutex> lk(g_tab_mu);
    for (auto &kv : g_tab) {
        cudaFree(kv.second.ptr);
    }
    g_tab.clear();
    g_tab_bytes = 0;
    g_tab_generation.fetch_add(1);
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
    if (n_in == 0) {  // the zero polynomial: nothing to read (d_in may be null)
        B2S_CUDA(cudaMemset2DAsync(d_out, sizeof(u64) * out_stride, 0, sizeof(u64) * n, n_planes, st));
        return 0;
    }
    const u64 w = inverse ? gl_inv(omega) : omega;        // code/ntt.py:41
    const u64 scale = inverse ? gl_inv(offset) : offset;  // code/ntt.py:165 / :174
    const bool do_scale = offset != 1;

    if (log_n <= 3) {
        const u64 ninv = gl_inv(n % GL_P);  // code/ntt.py:39
        small_dft_kernel<<<n_planes, 32, 0, st>>>(d_in, in_stride, n_in, d_out, out_stride, (u32)n, w,
                                                  inverse ? 1 : scale, inverse ? ninv : 1, inverse ? scale : 1);
        B2S_LAUNCHED();
        return 0;
    }

    // plans (digit split, strides, constant twiddles, table pointers) are cached per transform
    // shape: building one costs ~15 us of host arithmetic, as much as a small transform on the device
    Pass4Plan plan[3];
    int npass = 0;
    // 8-point core steps give a small job twice the threads and half the per-thread chain: faster
    // up to 2^19 elements in total (2^18: 14.5 vs 18.5 us); 16-point steps need ~20 % fewer
    // instructions and win once the GPU is full (profiles/r01_ntt_experiments.md)
    static const char *force_e = getenv("B2S_NTT_LOG_E");
    const u32 log_E = force_e ? (u32)atoi(force_e) : ((u64)n * n_planes <= ((u64)1 << 19) ? 3 : 4);
    int rc = get_plan(log_n, n_in, w, scale, inverse != 0, do_scale, log_E, st, plan, &npass);
    if (rc) return rc;
    u64 *work = nullptr;
    if (npass > 1) {
        rc = get_work(sizeof(u64) * n * n_planes, st, &work);
        if (rc) return rc;
    }
    for (int ps = 0; ps < npass && rc == 0; ++ps) {
        Pass4Plan &pl = plan[ps];
        pl.P.in = pl.first ? d_in : work;
        pl.P.in_plane_stride = pl.first ? in_stride : n;
        pl.P.out = pl.last ? d_out : work;
        pl.P.out_plane_stride = pl.last ? out_stride : n;
        rc = dispatch_pass4(pl, n_planes, st);
    }
    return rc;
}
