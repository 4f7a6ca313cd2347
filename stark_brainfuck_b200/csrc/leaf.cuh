// leaf.cuh -- device-side emitter of pickle.dumps(field element) and its BLAKE2b digest.
//
// code/merkle.py:29-32 hashes pickle.dumps(leaf).  For field-element leaves the pickle is
// a fixed byte template (derived on the host from the caller's own classes, see
// include/b2s.h) with the coefficient integers spliced in; CPython 3.12 protocol-4
// integer encoding (Modules/_pickle.c save_long):
//     v < 2^8  : 'K' v            v < 2^16 : 'M' v(2 LE)        v < 2^31 : 'J' v(4 LE)
//     else     : 0x8a nbytes v(nbytes LE), nbytes = (bit_length(v) >> 3) + 1
// and the frame header 80 04 95 <u64 LE body length>.
//
// Two things keep the hash count down.  (1) Everything before the first integer is constant
// up to the frame length, so the BLAKE2b state after the leading full blocks depends only on
// the template k and the total length L: the host precomputes those midstates (LeafMid) and a
// thread starts from mid[k][L - lmin[k]] -- 3 compressions instead of 4 for a random
// extension-field leaf.  (2) Only the bytes after that cut are materialised, in a per-thread
// shared-memory slot, with word-wide copies of the template segments.
#pragma once
#include "blake2b.cuh"
#include "common.h"

#define LEAF_MAXD 28  // total length varies by at most 9 bytes per integer, 3 integers

// host-prepared companion of b2s_leaf_templates (merkle.cu prepare_leaf)
struct LeafMid {
    u32 cut[4];         // bytes covered by the midstate of template k (multiple of 128)
    u32 lmin[4];        // total length of template k with 2-byte integers
    const u64 *mid;     // device: [k][d][8] BLAKE2b state after the first cut[k] bytes, d = L - lmin[k]
};

// shared-memory copy of the templates (byte addressed with divergent offsets per lane)
struct LeafTplSmem {
    u32 seg_off[4][5];
    u32 cut[4], lmin[4];
    u8 bytes[B2S_TPL_MAX_BYTES + 8];  // + slack: word-wide copies read up to 7 bytes past a segment
};

__device__ __forceinline__ void leaf_tpl_to_smem(const b2s_leaf_templates &tpl, const LeafMid &lm, LeafTplSmem *s) {
    const u32 *src = reinterpret_cast<const u32 *>(tpl.bytes);
    u32 *dst = reinterpret_cast<u32 *>(s->bytes);
    for (int i = threadIdx.x; i < B2S_TPL_MAX_BYTES / 4; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x < 2) dst[B2S_TPL_MAX_BYTES / 4 + threadIdx.x] = 0;
    if (threadIdx.x < 20) (&s->seg_off[0][0])[threadIdx.x] = (&tpl.seg_off[0][0])[threadIdx.x];
    if (threadIdx.x < 4) {
        s->cut[threadIdx.x] = lm.cut[threadIdx.x];
        s->lmin[threadIdx.x] = lm.lmin[threadIdx.x];
    }
}

// per-thread slot for the part of the preimage behind the midstate cut: MB blocks of 128 bytes
template <int MB>
struct LeafCfg {
    static constexpr int MAX_MSG = MB * 128;
    static constexpr int MSG_STRIDE = MAX_MSG + 8;  // 8 * odd: conflict-free 64-bit reads across lanes
};

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)(uintptr_t)p; }

// length of the pickled integer (opcode + payload)
__device__ __forceinline__ u32 pickle_int_len(u64 v) {
    if (v < 256) return 2;
    if (v < 65536) return 3;
    if (v < 0x80000000ULL) return 5;
    return 3 + ((64 - __clzll((long long)v)) >> 3);
}

// shared -> shared byte copy with 32-bit stores (source and destination arbitrarily aligned)
__device__ __forceinline__ void leaf_copy(u8 *dst, const u8 *src, u32 n) {
    while (n && (smem_addr(dst) & 3)) {
        *dst++ = *src++;
        --n;
    }
    if (n >= 4) {
        const u32 sh = (smem_addr(src) & 3) * 8;
        const u32 *sw = reinterpret_cast<const u32 *>(src - (smem_addr(src) & 3));
        u32 lo = sw[0];
        do {
            const u32 hi = sw[1];
            *reinterpret_cast<u32 *>(dst) = __funnelshift_r(lo, hi, sh);
            lo = hi;
            ++sw;
            dst += 4;
            src += 4;
            n -= 4;
        } while (n >= 4);
    }
    while (n) {
        *dst++ = *src++;
        --n;
    }
}

// Digest of one field-element leaf.  `msg` is the thread's shared-memory slot (MSG_STRIDE bytes).
template <int NSLOTS>
__device__ __forceinline__ void leaf_digest(const u64 (&c)[3], bool trim, const LeafTplSmem *tp, const u64 *mid, u8 *msg,
                                            u64 h[8]) {
    u64 *msg64 = reinterpret_cast<u64 *>(msg);
    int k = NSLOTS;
    if (trim) {
        while (k > 0 && c[k - 1] == 0) --k;
    }
    u32 ilen[3] = {0, 0, 0};
    u32 len = tp->lmin[k] - 2 * k;
#pragma unroll
    for (int j = 0; j < NSLOTS; ++j) {
        if (j < k) {
            ilen[j] = pickle_int_len(c[j]);
            len += ilen[j];
        }
    }
    const u32 cut = tp->cut[k];
    const u32 first = cut >> 7, nblocks = (len + 127) >> 7;  // first < nblocks by construction
    {   // start from the precomputed state (or the IV when nothing is constant)
        const ulonglong2 *m2 = reinterpret_cast<const ulonglong2 *>(mid + ((u32)k * LEAF_MAXD + (len - tp->lmin[k])) * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const ulonglong2 v = m2[i];
            h[2 * i] = v.x;
            h[2 * i + 1] = v.y;
        }
    }
    // the final block is only partly overwritten: clear it first
    u64 *lastblk = msg64 + (nblocks - 1 - first) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) lastblk[i] = 0;

    u32 p = 11;  // logical position in the preimage; byte q lives at msg[q - cut]
    if (cut == 0) {
        const u64 body = len - 11;
        msg[0] = 0x80;
        msg[1] = 0x04;
        msg[2] = 0x95;
#pragma unroll
        for (int i = 0; i < 8; ++i) msg[3 + i] = (u8)(body >> (8 * i));
    }
    for (int j = 0; j <= k; ++j) {
        const u32 a = tp->seg_off[k][j], n = tp->seg_off[k][j + 1] - a;
        if (p + n > cut) {
            const u32 skip = p < cut ? cut - p : 0;
            leaf_copy(msg + (p + skip - cut), tp->bytes + a + skip, n - skip);
        }
        p += n;
        if (j < k) {
            const u64 v = c[j];
            u8 *o = msg + (p - cut);
            if (v < 256) {
                o[0] = 0x4b;
                o[1] = (u8)v;
            } else if (v < 65536) {
                o[0] = 0x4d;
                o[1] = (u8)v;
                o[2] = (u8)(v >> 8);
            } else if (v < 0x80000000ULL) {
                o[0] = 0x4a;
                for (int i = 0; i < 4; ++i) o[1 + i] = (u8)(v >> (8 * i));
            } else {
                const u32 nb = ilen[j] - 2;
                o[0] = 0x8a;
                o[1] = (u8)nb;
                for (u32 i = 0; i < nb; ++i) o[2 + i] = i < 8 ? (u8)(v >> (8 * i)) : 0;
            }
            p += ilen[j];
        }
    }
    for (u32 blk = first; blk < nblocks; ++blk) {
        u64 m[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = msg64[(blk - first) * 16 + i];
        const bool last = blk + 1 == nblocks;
        b2b_compress(h, m, last ? (u64)len : (u64)(blk + 1) * 128, last);
    }
}

__device__ __forceinline__ void store_digest(u8 *nodes, u64 slot, const u64 h[8]) {
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(nodes + slot * 64);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_ulonglong2(h[2 * i], h[2 * i + 1]);
}

// parent = blake2b(left | right): exactly one final 128-byte block (code/merkle.py:38-39)
__device__ __forceinline__ void node_digest(const u8 *nodes, u64 k, u64 h[8]) {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(nodes + k * 128);
    u64 m[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ulonglong2 v = src[i];
        m[2 * i] = v.x;
        m[2 * i + 1] = v.y;
    }
    b2b_init(h);
    b2b_compress(h, m, 128, true);
}

// ---- in-CTA subtree: digests live in shared memory word-major, D[w * cnt + i] = word w of node i,
// so that a parent reads (left, right) words as one 16-byte access and lanes stay contiguous.
// Reduces `cnt` nodes (a power of two >= 2, nodes heap-indexed heap0 .. heap0+cnt-1) to one,
// writing every level to the heap.  All threads of the CTA must call it.
__device__ __forceinline__ void subtree_reduce(u64 *D, u64 *D2, u32 cnt, u64 heap0, u8 *nodes) {
    u64 *src = D, *dst = D2;
    while (cnt > 1) {
        const u32 half = cnt >> 1;
        heap0 >>= 1;
        __syncthreads();
        for (u32 j = threadIdx.x; j < half; j += blockDim.x) {
            u64 m[16], h[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + w * cnt + 2 * j);
                m[w] = v.x;
                m[8 + w] = v.y;
            }
            b2b_init(h);
            b2b_compress(h, m, 128, true);
#pragma unroll
            for (int w = 0; w < 8; ++w) dst[w * half + j] = h[w];
            store_digest(nodes, heap0 + j, h);
        }
        u64 *t = src;
        src = dst;
        dst = t;
        cnt = half;
    }
}
