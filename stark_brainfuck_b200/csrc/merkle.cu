// merkle.cu -- BLAKE2b-512 Merkle trees in the reference's heap layout (code/merkle.py:8-52)
// and one FRI commit round (code/fri.py:127-128) fused with the Merkle tree of the folded
// codeword (code/fri.py:108 of the next round).
//
// Launch structure: ONE kernel hashes 128 leaves per CTA (optionally folding them first) and
// reduces them to the root of their 128-leaf subtree in shared memory, writing all 255 nodes;
// a second kernel reduces 256 nodes per CTA by 8 levels.  A 2^20-leaf tree is 3 launches
// (8192 + 32 + 1 CTAs) instead of one launch per level.
#include <string.h>

#include <mutex>
#include <vector>

#include "leaf.cuh"

int merkle_upper_run(u8 *d_nodes, u64 npo2, cudaStream_t st);

namespace {

constexpr int LEAF_THREADS = 128;    // leaves per CTA of the leaf kernels
constexpr int REDUCE_NODES = 256;    // nodes per CTA of the reduce kernel

struct FoldParams {
    const u64 *cw;
    u64 *next;
    u64 cw_stride, next_stride;
    u64 alpha[3];
    u64 inv_offset;
    u64 winv_sq[32];  // (omega^-1)^(2^b)
};

template <int MB>
constexpr size_t leaf_smem_bytes() {
    return ((sizeof(LeafTplSmem) + 15) & ~(size_t)15) + (size_t)LEAF_THREADS * LeafCfg<MB>::MSG_STRIDE;
}

// Leaves = field elements given as planes (FOLD = false) or the fold of codeword `F.cw` (FOLD =
// true, which also writes the folded planes).  n = number of leaves (a power of two).
// SUBTREE: also reduce the CTA's leaves to their subtree root in shared memory (small trees:
// fewer launches; the low-parallelism upper levels would cost occupancy on large ones).
template <int NSLOTS, int MB, bool FOLD, bool SUBTREE>
__global__ void __launch_bounds__(LEAF_THREADS)
    leaf_subtree_kernel(const u64 *__restrict__ planes, u64 stride, u64 n, const __grid_constant__ b2s_leaf_templates tpl,
                        const __grid_constant__ LeafMid lm, const __grid_constant__ FoldParams F, u8 *__restrict__ nodes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LeafTplSmem *tp = reinterpret_cast<LeafTplSmem *>(smem_raw);
    u8 *msgs = smem_raw + ((sizeof(LeafTplSmem) + 15) & ~(size_t)15);
    leaf_tpl_to_smem(tpl, lm, tp);
    __syncthreads();
    pdl_wait();  // the codeword (the previous round's fold) may still be in flight
    pdl_trigger();
    const u64 base = (u64)blockIdx.x * LEAF_THREADS;
    const u64 i = base + threadIdx.x;
    if (nodes && i < 4) reinterpret_cast<ulonglong2 *>(nodes)[i] = make_ulonglong2(0, 0);  // slot 0 is never a node
    const u32 cnt = n < LEAF_THREADS ? (u32)n : LEAF_THREADS;  // leaves of this CTA
    u64 h[8];
    if (threadIdx.x < cnt) {
        u64 c[3] = {0, 0, 0};
        if (FOLD) {
            // 2^-1 * ((1 + alpha/x) a + (1 - alpha/x) b) = (a + b)/2 + alpha * ((a - b) / (2x)),  x = offset * omega^i
            const u64 xinv = gl_mul(F.inv_offset, gl_pow_sq(F.winv_sq, i));
            xfe a, b;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                a.c[s] = F.cw[s * F.cw_stride + i];
                b.c[s] = F.cw[s * F.cw_stride + n + i];
            }
            const xfe sum = x_mul_base(x_add(a, b), GL_HALF);
            const xfe dif = x_mul_base(x_sub(a, b), gl_mul(GL_HALF, xinv));
            const xfe al = {{F.alpha[0], F.alpha[1], F.alpha[2]}};
            const xfe r = x_add(sum, x_mul(al, dif));
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                c[s] = r.c[s];
                F.next[s * F.next_stride + i] = r.c[s];
            }
        } else {
#pragma unroll
            for (int s = 0; s < NSLOTS; ++s) c[s] = planes[s * stride + i];
        }
        if (nodes) {
            leaf_digest<NSLOTS>(c, tpl.trim != 0, tp, lm.mid, msgs + threadIdx.x * LeafCfg<MB>::MSG_STRIDE, h);
            store_digest(nodes, n + i, h);
        }
    }
    if (!SUBTREE || !nodes || cnt < 2) return;
    __syncthreads();  // message slots are dead: reuse them for the digests of the subtree
    u64 *D = reinterpret_cast<u64 *>(msgs);
    u64 *D2 = D + 8 * LEAF_THREADS;
    if (threadIdx.x < cnt) {
#pragma unroll
        for (int w = 0; w < 8; ++w) D[w * cnt + threadIdx.x] = h[w];
    }
    subtree_reduce(D, D2, cnt, n + base, nodes);
}

// nodes[first .. first+count) from their children: one thread per node (large levels)
__global__ void __launch_bounds__(128) merkle_level_kernel(u8 *nodes, u64 first, u64 count) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (j >= count) return;
    u64 h[8];
    node_digest(nodes, first + j, h);
    store_digest(nodes, first + j, h);
}

// nodes [count, 2 count) are known; every CTA reduces REDUCE_NODES of them (or all `count`).
__global__ void __launch_bounds__(REDUCE_NODES / 2) merkle_reduce_kernel(u8 *nodes, u64 count) {
    __shared__ __align__(16) u64 D[8 * REDUCE_NODES / 2];
    __shared__ __align__(16) u64 D2[8 * REDUCE_NODES / 4];
    const u32 cnt = count < REDUCE_NODES ? (u32)count : REDUCE_NODES;
    const u64 heap0 = count + (u64)blockIdx.x * cnt;
    const u32 half = cnt >> 1;
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x < half) {
        u64 h[8];
        node_digest(nodes, (heap0 >> 1) + threadIdx.x, h);
        store_digest(nodes, (heap0 >> 1) + threadIdx.x, h);
#pragma unroll
        for (int w = 0; w < 8; ++w) D[w * half + threadIdx.x] = h[w];
    }
    if (half >= 2) subtree_reduce(D, D2, half, heap0 >> 1, nodes);
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
#pragma unroll 1
    for (u64 blk = 0; blk < nblocks; ++blk) {
        // Full blocks and the bytes of the last one are read with plain byte loads under ordinary branches.  (The
        // obvious `q < len ? msg[q] : 0` inside the unrolled loops compiled -- nvcc 12.9, sm_100a -- into predicated
        // byte loads whose destination aliases the address register and is not cleared when the predicate is off:
        // profiles/microbench/blob_san_repro.cu hashes a 427-byte message wrongly with it.)
        u64 m[16];
        const u64 base = blk * 128;
        const u64 have = len - base < 128 ? len - base : 128;  // bytes of this block
        if (have == 128) {
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                u64 v = 0;
#pragma unroll
                for (int b = 7; b >= 0; --b) v = (v << 8) | msg[base + w * 8 + b];
                m[w] = v;
            }
        } else {
#pragma unroll
            for (int w = 0; w < 16; ++w) m[w] = 0;
            for (u64 q = 0; q < have; ++q) {
                const u64 byte = msg[base + q];
                const u32 w = (u32)(q >> 3), sh = (u32)(q & 7) * 8;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k == (int)w) m[k] |= byte << sh;  // static register indices
            }
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

// ---- host: BLAKE2b midstates of the constant leaf prefix ------------------------------------------
inline u64 rotr64(u64 x, int r) { return (x >> r) | (x << (64 - r)); }

void host_b2b_compress(u64 h[8], const u8 block[128], u64 t, bool last) {
    static const u64 IV[8] = {B2B_IV0, B2B_IV1, B2B_IV2, B2B_IV3, B2B_IV4, B2B_IV5, B2B_IV6, B2B_IV7};
    static const u8 SIGMA[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    u64 m[16], v[16];
    memcpy(m, block, 128);  // little-endian host
    for (int i = 0; i < 8; ++i) {
        v[i] = h[i];
        v[8 + i] = IV[i];
    }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
    auto G = [&](int a, int b, int c, int d, u64 x, u64 y) {
        v[a] = v[a] + v[b] + x;
        v[d] = rotr64(v[d] ^ v[a], 32);
        v[c] = v[c] + v[d];
        v[b] = rotr64(v[b] ^ v[c], 24);
        v[a] = v[a] + v[b] + y;
        v[d] = rotr64(v[d] ^ v[a], 16);
        v[c] = v[c] + v[d];
        v[b] = rotr64(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; ++r) {
        const u8 *s = SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[8 + i];
}

struct PreparedLeaf {
    b2s_leaf_templates tpl;
    LeafMid lm;
    u32 msg_blocks;  // blocks a thread has to materialise at most
    int dev;
};
std::mutex g_leaf_mu;
std::vector<PreparedLeaf *> g_leaf_cache;

int check_templates(const b2s_leaf_templates *tpl) {
    if (!tpl || (tpl->n_slots != 1 && tpl->n_slots != 3)) {
        b2s_set_error("leaf templates: n_slots must be 1 or 3");
        return B2S_ERR_ARG;
    }
    for (u32 k = tpl->trim ? 0 : tpl->n_slots; k <= tpl->n_slots; ++k)
        for (u32 j = 0; j <= k; ++j)
            if (tpl->seg_off[k][j + 1] < tpl->seg_off[k][j] || tpl->seg_off[k][j + 1] > B2S_TPL_MAX_BYTES) {
                b2s_set_error("leaf templates: bad segment offsets");
                return B2S_ERR_ARG;
            }
    return 0;
}

// midstates + cut positions for a template set; cached per (template bytes, device)
int prepare_leaf(const b2s_leaf_templates *tpl, const PreparedLeaf **out) {
    int rc = check_templates(tpl);
    if (rc) return rc;
    int dev = 0;
    B2S_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_leaf_mu);
    for (PreparedLeaf *e : g_leaf_cache)
        if (e->dev == dev && e->tpl.n_slots == tpl->n_slots && e->tpl.trim == tpl->trim &&
            !memcmp(e->tpl.seg_off, tpl->seg_off, sizeof(tpl->seg_off)) &&
            !memcmp(e->tpl.bytes, tpl->bytes, sizeof(tpl->bytes))) {
            *out = e;
            return 0;
        }
    PreparedLeaf *e = new PreparedLeaf;
    e->tpl = *tpl;
    e->dev = dev;
    e->msg_blocks = 1;
    std::vector<u64> mid((size_t)4 * LEAF_MAXD * 8, 0);
    for (u32 k = 0; k < 4; ++k) e->lm.cut[k] = e->lm.lmin[k] = 0;
    for (u32 k = tpl->trim ? 0 : tpl->n_slots; k <= tpl->n_slots; ++k) {
        u32 base = 11;
        for (u32 j = 0; j <= k; ++j) base += tpl->seg_off[k][j + 1] - tpl->seg_off[k][j];
        const u32 lmin = base + 2 * k, lmax = base + 11 * k;
        const u32 seg0 = tpl->seg_off[k][1] - tpl->seg_off[k][0];
        const u32 cp = 11 + seg0;  // bytes in front of the first integer (k == 0: the whole preimage)
        u32 nconst = cp / 128;
        if (k == 0) nconst = (lmin + 127) / 128 - 1;
        e->lm.cut[k] = nconst * 128;
        e->lm.lmin[k] = lmin;
        const u32 need = (lmax - nconst * 128 + 127) / 128;
        if (need > e->msg_blocks) e->msg_blocks = need;
        for (u32 d = 0; d <= lmax - lmin && d < LEAF_MAXD; ++d) {
            u64 h[8] = {B2B_IV0 ^ 0x01010040ULL, B2B_IV1, B2B_IV2, B2B_IV3, B2B_IV4, B2B_IV5, B2B_IV6, B2B_IV7};
            std::vector<u8> pre(nconst * 128 + 128, 0);
            const u64 body = lmin + d - 11;
            pre[0] = 0x80;
            pre[1] = 0x04;
            pre[2] = 0x95;
            memcpy(&pre[3], &body, 8);
            const u32 take = nconst * 128 > 11 ? nconst * 128 - 11 : 0;  // <= seg0 by construction
            memcpy(&pre[11], tpl->bytes + tpl->seg_off[k][0], take);
            for (u32 b = 0; b < nconst; ++b) host_b2b_compress(h, &pre[b * 128], (u64)(b + 1) * 128, false);
            memcpy(&mid[((size_t)k * LEAF_MAXD + d) * 8], h, 64);
        }
    }
    const u32 cap = tpl->n_slots == 3 ? 4 : 2;
    if (e->msg_blocks > cap) {
        b2s_set_error("leaf templates: preimage tail of %u blocks exceeds the device buffer (%u)", e->msg_blocks, cap);
        delete e;
        return B2S_ERR_ARG;
    }
    u64 *d_mid = nullptr;
    B2S_CUDA(cudaMalloc(&d_mid, mid.size() * sizeof(u64)));
    B2S_CUDA(cudaMemcpy(d_mid, mid.data(), mid.size() * sizeof(u64), cudaMemcpyHostToDevice));
    e->lm.mid = d_mid;
    g_leaf_cache.push_back(e);
    *out = e;
    return 0;
}

constexpr u64 SUBTREE_MAX_LEAVES = 8192;  // up to here the leaf kernel also builds its 128-leaf subtrees
constexpr u64 LEVEL_MIN_NODES = 4096;     // levels with more nodes than this get one launch each

template <int NSLOTS, int MB, bool FOLD, bool SUBTREE>
int launch_leaf2(const u64 *d_planes, u64 stride, u64 n, const PreparedLeaf *pl, const FoldParams &F, u8 *d_nodes,
                 cudaStream_t st) {
    static bool attr[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    constexpr size_t smem = leaf_smem_bytes<MB>();
    if (!attr[dev & 15]) {
        B2S_CUDA(cudaFuncSetAttribute(leaf_subtree_kernel<NSLOTS, MB, FOLD, SUBTREE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[dev & 15] = true;
    }
    const unsigned blocks = (unsigned)((n + LEAF_THREADS - 1) / LEAF_THREADS);
    B2S_CUDA(launch_pdl(leaf_subtree_kernel<NSLOTS, MB, FOLD, SUBTREE>, dim3(blocks), dim3(LEAF_THREADS), smem, st, d_planes,
                        stride, n, pl->tpl, pl->lm, F, d_nodes));
    B2S_LAUNCHED();
    return 0;
}

// leaf digests (+ fold), then the rest of the tree
template <int NSLOTS, int MB, bool FOLD>
int launch_leaf(const u64 *d_planes, u64 stride, u64 n, const PreparedLeaf *pl, const FoldParams &F, u8 *d_nodes,
                cudaStream_t st) {
    if (!d_nodes || n > SUBTREE_MAX_LEAVES) {
        int rc = launch_leaf2<NSLOTS, MB, FOLD, false>(d_planes, stride, n, pl, F, d_nodes, st);
        return rc || !d_nodes ? rc : merkle_upper_run(d_nodes, n, st);
    }
    int rc = launch_leaf2<NSLOTS, MB, FOLD, true>(d_planes, stride, n, pl, F, d_nodes, st);
    return rc || n <= LEAF_THREADS ? rc : merkle_upper_run(d_nodes, n / LEAF_THREADS, st);
}

}  // namespace

// all levels above `npo2` known nodes [npo2, 2 npo2)
int merkle_upper_run(u8 *d_nodes, u64 npo2, cudaStream_t st) {
    u64 cnt = npo2;
    while (cnt > LEVEL_MIN_NODES) {  // large levels: full parallelism, one launch each
        cnt >>= 1;
        B2S_CUDA(launch_pdl(merkle_level_kernel, dim3((unsigned)((cnt + 127) / 128)), dim3(128), 0, st, d_nodes, cnt, cnt));
        B2S_LAUNCHED();
    }
    while (cnt > 1) {  // the top: 8 levels per launch
        const u64 per = cnt < REDUCE_NODES ? cnt : REDUCE_NODES;
        B2S_CUDA(launch_pdl(merkle_reduce_kernel, dim3((unsigned)(cnt / per)), dim3(REDUCE_NODES / 2), 0, st, d_nodes, cnt));
        B2S_LAUNCHED();
        cnt /= per;
    }
    return 0;
}

int merkle_field_run(const u64 *d_planes, u64 stride, u64 n, const b2s_leaf_templates *tpl, u8 *d_nodes,
                     cudaStream_t st) {
    if (n == 0 || (n & (n - 1))) {
        b2s_set_error("field-element Merkle trees need a power-of-two leaf count, got %llu", (unsigned long long)n);
        return B2S_ERR_ARG;
    }
    const PreparedLeaf *pl = nullptr;
    int rc = prepare_leaf(tpl, &pl);
    if (rc) return rc;
    FoldParams F = {};
    if (tpl->n_slots == 3)
        rc = pl->msg_blocks <= 3 ? launch_leaf<3, 3, false>(d_planes, stride, n, pl, F, d_nodes, st)
                                 : launch_leaf<3, 4, false>(d_planes, stride, n, pl, F, d_nodes, st);
    else
        rc = launch_leaf<1, 2, false>(d_planes, stride, n, pl, F, d_nodes, st);
    return rc;
}

extern "C" int b2s_merkle_field(const uint64_t *d_planes, uint64_t plane_stride, uint64_t n,
                                const b2s_leaf_templates *tpl, uint8_t *d_nodes, void *stream) {
    return merkle_field_run(d_planes, plane_stride, n, tpl, d_nodes, (cudaStream_t)stream);
}

extern "C" int b2s_merkle_upper(uint8_t *d_nodes, uint64_t npo2, void *stream) {
    if (npo2 == 0 || (npo2 & (npo2 - 1))) {
        b2s_set_error("merkle_upper: node count must be a power of two, got %llu", (unsigned long long)npo2);
        return B2S_ERR_ARG;
    }
    return merkle_upper_run(d_nodes, npo2, (cudaStream_t)stream);
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
    FoldParams F;
    F.cw = d_cw;
    F.next = d_next;
    F.cw_stride = cw_stride;
    F.next_stride = next_stride;
    for (int i = 0; i < 3; ++i) F.alpha[i] = alpha[i];
    F.inv_offset = gl_inv(offset);
    u64 sq = gl_inv(omega);
    for (int b = 0; b < 32; ++b) {
        F.winv_sq[b] = sq;
        sq = gl_mul(sq, sq);
    }
    const u64 half = N / 2;
    static b2s_leaf_templates dummy = [] {
        b2s_leaf_templates t;
        memset(&t, 0, sizeof(t));
        t.n_slots = 3;
        t.trim = 1;
        return t;
    }();
    const b2s_leaf_templates *use = d_next_nodes ? tpl : &dummy;
    if (d_next_nodes && (!tpl || tpl->n_slots != 3)) {
        b2s_set_error("fri_fold: extension-field leaf templates required");
        return B2S_ERR_ARG;
    }
    const PreparedLeaf *pl = nullptr;
    int rc = prepare_leaf(use, &pl);
    if (rc) return rc;
    return pl->msg_blocks <= 3 ? launch_leaf<3, 3, true>(nullptr, 0, half, pl, F, d_next_nodes, st)
                               : launch_leaf<3, 4, true>(nullptr, 0, half, pl, F, d_next_nodes, st);
}
