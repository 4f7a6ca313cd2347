// leaf.cuh -- device-side emitter of pickle.dumps(field element) and its BLAKE2b digest.
//
// code/merkle.py:29-32 hashes pickle.dumps(leaf).  For field-element leaves the pickle is
// a fixed byte template (derived on the host from the caller's own classes, see
// include/b2s.h) with the coefficient integers spliced in; CPython 3.12 protocol-4
// integer encoding (Modules/_pickle.c save_long):
//     v < 2^8  : 'K' v            v < 2^16 : 'M' v(2 LE)        v < 2^31 : 'J' v(4 LE)
//     else     : 0x8a nbytes v(nbytes LE), nbytes = (bit_length(v) >> 3) + 1
// and the frame header 80 04 95 <u64 LE body length>.
#pragma once
#include "blake2b.cuh"
#include "common.h"

// shared-memory copy of the templates (byte addressed with divergent offsets per lane)
struct LeafTplSmem {
    u32 seg_off[4][5];
    u8 bytes[B2S_TPL_MAX_BYTES];
};

__device__ __forceinline__ void leaf_tpl_to_smem(const b2s_leaf_templates &tpl, LeafTplSmem *s) {
    const u32 *src = reinterpret_cast<const u32 *>(tpl.bytes);
    u32 *dst = reinterpret_cast<u32 *>(s->bytes);
    for (int i = threadIdx.x; i < B2S_TPL_MAX_BYTES / 4; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x < 20) (&s->seg_off[0][0])[threadIdx.x] = (&tpl.seg_off[0][0])[threadIdx.x];
}

template <int NSLOTS>
struct LeafCfg {
    static constexpr int MAX_MSG = NSLOTS == 3 ? 512 : 256;  // bytes, multiple of 128
    static constexpr int MSG_STRIDE = MAX_MSG + 8;           // per-thread slot (keeps u64 alignment, skews banks)
};

// Builds the preimage of one leaf in `msg` (per-thread shared-memory slot) and returns its
// BLAKE2b-512 digest in h[8].
template <int NSLOTS>
__device__ __forceinline__ void leaf_digest(const u64 (&c)[3], bool trim, const LeafTplSmem *tp, u8 *msg, u64 h[8]) {
    constexpr int MAX_MSG = LeafCfg<NSLOTS>::MAX_MSG;
    u64 *msg64 = reinterpret_cast<u64 *>(msg);
#pragma unroll 4
    for (int i = 0; i < MAX_MSG / 8; ++i) msg64[i] = 0;
    int k = NSLOTS;
    if (trim) {
        while (k > 0 && c[k - 1] == 0) --k;
    }
    u32 p = 11;
    for (int j = 0; j <= k; ++j) {
        const u32 a = tp->seg_off[k][j], b = tp->seg_off[k][j + 1];
        for (u32 q = a; q < b; ++q) msg[p++] = tp->bytes[q];
        if (j < k) {
            const u64 v = c[j];
            if (v < 256) {
                msg[p++] = 0x4b;
                msg[p++] = (u8)v;
            } else if (v < 65536) {
                msg[p++] = 0x4d;
                msg[p++] = (u8)v;
                msg[p++] = (u8)(v >> 8);
            } else if (v < 0x80000000ULL) {
                msg[p++] = 0x4a;
                for (int i = 0; i < 4; ++i) msg[p++] = (u8)(v >> (8 * i));
            } else {
                const u32 nb = ((64 - __clzll((long long)v)) >> 3) + 1;
                msg[p++] = 0x8a;
                msg[p++] = (u8)nb;
                for (u32 i = 0; i < nb; ++i) msg[p++] = i < 8 ? (u8)(v >> (8 * i)) : 0;
            }
        }
    }
    const u64 body = p - 11;
    msg[0] = 0x80;
    msg[1] = 0x04;
    msg[2] = 0x95;
    for (int i = 0; i < 8; ++i) msg[3 + i] = (u8)(body >> (8 * i));

    b2b_init(h);
    const u32 len = p;
    const u32 nblocks = (len + 127) >> 7;  // len >= 11
    for (u32 blk = 0; blk < nblocks; ++blk) {
        u64 m[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = msg64[blk * 16 + i];
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
