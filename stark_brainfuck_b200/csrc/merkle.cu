// merkle.cu -- BLAKE2b-512 Merkle trees in the reference's heap layout (code/merkle.py:8-52)
// and one FRI commit round (code/fri.py:127-128) fused with the leaf hashing of the folded
// codeword (code/fri.py:108 of the next round).
#include "leaf.cuh"

namespace {

constexpr int LEAF_THREADS = 64;

template <int NSLOTS>
__global__ void __launch_bounds__(LEAF_THREADS)
    merkle_leaf_kernel(const u64 *__restrict__ planes, u64 stride, u64 n, const __grid_constant__ b2s_leaf_templates tpl,
                       u8 *__restrict__ nodes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LeafTplSmem *tp = reinterpret_cast<LeafTplSmem *>(smem_raw);
    u8 *msgs = smem_raw + ((sizeof(LeafTplSmem) + 15) & ~(size_t)15);
    leaf_tpl_to_smem(tpl, tp);
    __syncthreads();
    const u64 i = (u64)blockIdx.x * LEAF_THREADS + threadIdx.x;
    if (i >= n) return;
    u64 c[3] = {0, 0, 0};
#pragma unroll
    for (int s = 0; s < NSLOTS; ++s) c[s] = planes[s * stride + i];
    u64 h[8];
    leaf_digest<NSLOTS>(c, tpl.trim != 0, tp, msgs + threadIdx.x * LeafCfg<NSLOTS>::MSG_STRIDE, h);
    store_digest(nodes, n + i, h);
}

// nodes[first .. first+count) from their children
__global__ void __launch_bounds__(128) merkle_level_kernel(u8 *nodes, u64 first, u64 count) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    u64 h[8];
    node_digest(nodes, first + j, h);
    store_digest(nodes, first + j, h);
}

// all levels with <= 512 nodes in one CTA (levels are dependent; block barrier between them)
__global__ void __launch_bounds__(512) merkle_top_kernel(u8 *nodes, u32 start_count) {
    for (u32 cnt = start_count; cnt >= 1; cnt >>= 1) {
        if (threadIdx.x < cnt) {
            u64 h[8];
            node_digest(nodes, (u64)cnt + threadIdx.x, h);
            store_digest(nodes, (u64)cnt + threadIdx.x, h);
        }
        __syncthreads();
    }
}

// generic leaves pickled by the host: digest of bytes[off[i] .. off[i+1])
__global__ void __launch_bounds__(128)
    merkle_blob_leaf_kernel(const u8 *__restrict__ bytes, const u64 *__restrict__ off, u64 n, u64 npo2, u8 *nodes) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8 *msg = bytes + off[i];
    const u64 len = off[i + 1] - off[i];
    u64 h[8];
    b2b_init(h);
    const u64 nblocks = len == 0 ? 1 : (len + 127) >> 7;
    for (u64 blk = 0; blk < nblocks; ++blk) {
        u64 m[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            u64 v = 0;
            for (int b = 7; b >= 0; --b) {
                const u64 q = blk * 128 + w * 8 + b;
                v = (v << 8) | (q < len ? msg[q] : 0);
            }
            m[w] = v;
        }
        const bool last = blk + 1 == nblocks;
        b2b_compress(h, m, last ? len : (blk + 1) * 128, last);
    }
    store_digest(nodes, npo2 + i, h);
}

// first parent level above blob leaves: missing leaves are the reference's 32-byte zero
// placeholders (code/merkle.py:26), so a parent hashes 128, 96 or 64 bytes.
__global__ void __launch_bounds__(128) merkle_blob_level1_kernel(u8 *nodes, u64 npo2, u64 n_leafs) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npo2 / 2) return;
    const u64 k = npo2 / 2 + j;
    const bool has_l = 2 * j < n_leafs, has_r = 2 * j + 1 < n_leafs;
    const u64 *src = reinterpret_cast<const u64 *>(nodes + k * 128);
    u64 m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = 0;
    if (has_l) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = src[i];
    }
    if (has_r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[8 + i] = src[8 + i];
    }
    const u64 len = (has_l ? 64 : 32) + (has_r ? 64 : 32);
    u64 h[8];
    b2b_init(h);
    b2b_compress(h, m, len, true);
    store_digest(nodes, k, h);
}

// ---- FRI fold (code/fri.py:127-128), optionally fused with next-round leaf hashing ----
struct FoldParams {
    const u64 *cw;
    u64 *next;
    u8 *next_nodes;  // may be null
    u64 cw_stride, next_stride, N;
    u64 alpha[3];
    u64 inv_offset;
    u64 winv_sq[32];  // (omega^-1)^(2^b)
};

template <bool HASH>
__global__ void __launch_bounds__(LEAF_THREADS)
    fri_fold_kernel(const __grid_constant__ FoldParams P, const __grid_constant__ b2s_leaf_templates tpl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LeafTplSmem *tp = reinterpret_cast<LeafTplSmem *>(smem_raw);
    u8 *msgs = smem_raw + ((sizeof(LeafTplSmem) + 15) & ~(size_t)15);
    if (HASH) {
        leaf_tpl_to_smem(tpl, tp);
        __syncthreads();
    }
    const u64 half = P.N >> 1;
    const u64 i = (u64)blockIdx.x * LEAF_THREADS + threadIdx.x;
    if (i >= half) return;
    // 1 / (offset * omega^i)
    const u64 xinv = gl_mul(P.inv_offset, gl_pow_sq(P.winv_sq, i));
    xfe a, b;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        a.c[s] = P.cw[s * P.cw_stride + i];
        b.c[s] = P.cw[s * P.cw_stride + half + i];
    }
    // 2^-1 * ((1 + alpha/x) a + (1 - alpha/x) b) = (a + b)/2 + alpha * ((a - b) / (2x))
    const xfe sum = x_mul_base(x_add(a, b), GL_HALF);
    const xfe dif = x_mul_base(x_sub(a, b), gl_mul(GL_HALF, xinv));
    const xfe al = {{P.alpha[0], P.alpha[1], P.alpha[2]}};
    const xfe r = x_add(sum, x_mul(al, dif));
#pragma unroll
    for (int s = 0; s < 3; ++s) P.next[s * P.next_stride + i] = r.c[s];
    if (HASH) {
        u64 h[8];
        const u64 c[3] = {r.c[0], r.c[1], r.c[2]};
        leaf_digest<3>(c, tpl.trim != 0, tp, msgs + threadIdx.x * LeafCfg<3>::MSG_STRIDE, h);
        store_digest(P.next_nodes, half + i, h);
    }
}

template <int NSLOTS>
size_t leaf_smem() {
    return ((sizeof(LeafTplSmem) + 15) & ~(size_t)15) + (size_t)LEAF_THREADS * LeafCfg<NSLOTS>::MSG_STRIDE;
}

int check_templates(const b2s_leaf_templates *tpl) {
    if (!tpl || (tpl->n_slots != 1 && tpl->n_slots != 3)) {
        b2s_set_error("leaf templates: n_slots must be 1 or 3");
        return B2S_ERR_ARG;
    }
    const u32 max_msg = tpl->n_slots == 3 ? LeafCfg<3>::MAX_MSG : LeafCfg<1>::MAX_MSG;
    for (u32 k = tpl->trim ? 0 : tpl->n_slots; k <= tpl->n_slots; ++k) {
        u32 tot = 11 + 11 * k;
        for (u32 j = 0; j <= k; ++j) {
            if (tpl->seg_off[k][j + 1] < tpl->seg_off[k][j] || tpl->seg_off[k][j + 1] > B2S_TPL_MAX_BYTES) {
                b2s_set_error("leaf templates: bad segment offsets");
                return B2S_ERR_ARG;
            }
            tot += tpl->seg_off[k][j + 1] - tpl->seg_off[k][j];
        }
        if (tot > max_msg) {
            b2s_set_error("leaf templates: preimage of up to %u bytes exceeds the device buffer (%u)", tot, max_msg);
            return B2S_ERR_ARG;
        }
    }
    return 0;
}

}  // namespace

int merkle_upper_run(u8 *d_nodes, u64 npo2, cudaStream_t st) {
    // levels with count = npo2/2 ... 1 (node indices [count, 2*count))
    u64 cnt = npo2 >> 1;
    while (cnt > 512) {
        merkle_level_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(d_nodes, cnt, cnt);
        B2S_LAUNCHED();
        cnt >>= 1;
    }
    if (cnt >= 1) {
        merkle_top_kernel<<<1, 512, 0, st>>>(d_nodes, (u32)cnt);
        B2S_LAUNCHED();
    }
    return 0;
}

int merkle_field_run(const u64 *d_planes, u64 stride, u64 n, const b2s_leaf_templates *tpl, u8 *d_nodes,
                     cudaStream_t st) {
    if (n == 0 || (n & (n - 1))) {
        b2s_set_error("field-element Merkle trees need a power-of-two leaf count, got %llu", (unsigned long long)n);
        return B2S_ERR_ARG;
    }
    int rc = check_templates(tpl);
    if (rc) return rc;
    B2S_CUDA(cudaMemsetAsync(d_nodes, 0, 64, st));
    const unsigned blocks = (unsigned)((n + LEAF_THREADS - 1) / LEAF_THREADS);
    if (tpl->n_slots == 3) {
        static bool attr = false;
        if (!attr) {
            B2S_CUDA(cudaFuncSetAttribute(merkle_leaf_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)leaf_smem<3>()));
            attr = true;
        }
        merkle_leaf_kernel<3><<<blocks, LEAF_THREADS, leaf_smem<3>(), st>>>(d_planes, stride, n, *tpl, d_nodes);
    } else {
        merkle_leaf_kernel<1><<<blocks, LEAF_THREADS, leaf_smem<1>(), st>>>(d_planes, stride, n, *tpl, d_nodes);
    }
    B2S_LAUNCHED();
    return merkle_upper_run(d_nodes, n, st);
}

extern "C" int b2s_merkle_field(const uint64_t *d_planes, uint64_t plane_stride, uint64_t n,
                                const b2s_leaf_templates *tpl, uint8_t *d_nodes, void *stream) {
    return merkle_field_run(d_planes, plane_stride, n, tpl, d_nodes, (cudaStream_t)stream);
}

extern "C" int b2s_merkle_blobs(const uint8_t *d_bytes, const uint64_t *d_offsets, uint64_t n_leafs, uint64_t npo2,
                                uint8_t *d_nodes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_leafs == 0 || npo2 < n_leafs || (npo2 & (npo2 - 1))) {
        b2s_set_error("merkle_blobs: bad leaf count %llu / %llu", (unsigned long long)n_leafs, (unsigned long long)npo2);
        return B2S_ERR_ARG;
    }
    B2S_CUDA(cudaMemsetAsync(d_nodes, 0, 128 * npo2, st));
    merkle_blob_leaf_kernel<<<(unsigned)((n_leafs + 127) / 128), 128, 0, st>>>(d_bytes, d_offsets, n_leafs, npo2,
                                                                              d_nodes);
    B2S_LAUNCHED();
    if (npo2 == 1) return 0;
    merkle_blob_level1_kernel<<<(unsigned)((npo2 / 2 + 127) / 128), 128, 0, st>>>(d_nodes, npo2, n_leafs);
    B2S_LAUNCHED();
    return merkle_upper_run(d_nodes, npo2 / 2, st);
}

extern "C" int b2s_fri_fold(const uint64_t *d_cw, uint64_t cw_stride, uint64_t N, const uint64_t alpha[3],
                            uint64_t offset, uint64_t omega, uint64_t *d_next, uint64_t next_stride,
                            const b2s_leaf_templates *tpl, uint8_t *d_next_nodes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N < 2 || (N & (N - 1))) {
        b2s_set_error("fri_fold: codeword length must be a power of two >= 2");
        return B2S_ERR_ARG;
    }
    if (offset == 0 || omega == 0) {
        b2s_set_error("fri_fold: divide by zero");  // code/extension_field.py:84
        return B2S_ERR_ARG;
    }
    FoldParams P;
    P.cw = d_cw;
    P.next = d_next;
    P.next_nodes = d_next_nodes;
    P.cw_stride = cw_stride;
    P.next_stride = next_stride;
    P.N = N;
    for (int i = 0; i < 3; ++i) P.alpha[i] = alpha[i];
    P.inv_offset = gl_inv(offset);
    u64 sq = gl_inv(omega);
    for (int b = 0; b < 32; ++b) {
        P.winv_sq[b] = sq;
        sq = gl_mul(sq, sq);
    }
    const u64 half = N / 2;
    const unsigned blocks = (unsigned)((half + LEAF_THREADS - 1) / LEAF_THREADS);
    if (d_next_nodes) {
        int rc = check_templates(tpl);
        if (rc) return rc;
        if (tpl->n_slots != 3) {
            b2s_set_error("fri_fold: extension-field leaf templates required");
            return B2S_ERR_ARG;
        }
        static bool attr = false;
        if (!attr) {
            B2S_CUDA(cudaFuncSetAttribute(fri_fold_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)leaf_smem<3>()));
            attr = true;
        }
        B2S_CUDA(cudaMemsetAsync(d_next_nodes, 0, 64, st));
        fri_fold_kernel<true><<<blocks, LEAF_THREADS, leaf_smem<3>(), st>>>(P, *tpl);
        B2S_LAUNCHED();
        return merkle_upper_run(d_next_nodes, half, st);
    }
    b2s_leaf_templates dummy;
    dummy.n_slots = 3;
    dummy.trim = 1;
    fri_fold_kernel<false><<<blocks, LEAF_THREADS, 16, st>>>(P, dummy);
    B2S_LAUNCHED();
    return 0;
}
