// rows.cu -- Merkle leaves that are ROWS of several codewords: the zipped, salted leaves of
// BrainfuckStark.prove (code/brainfuck_stark.py:178-180, :197-199 -> code/salted_merkle.py:25-35).
//
// Leaf r = blake2b(pickle.dumps(row_r) | pickle.dumps(salt_r)), row_r = the tuple of the r-th elements of
// every codeword.  The elements of one row carry different `field` objects and the first occurrence of
// every class / string / field object defines a memo entry that later elements reference, so the pickle
// of a row is one byte template per proof -- derived on the host from the caller's own objects -- with
// one integer spliced in per coefficient:
//     PROTO 4 | FRAME(len) | seg_0 INT(v_0) seg_1 INT(v_1) ... seg_S | salt_prefix salt_r salt_suffix
// The integers are read from the codeword planes where the transforms left them.  A row whose shape
// differs from the template's (an extension-field element with trimmed coefficients,
// code/extension_field.py:6-9) is not hashed; its index is reported and the caller hashes those rows with
// the template of their own shape (row list).
//
// One thread per row.  The preimage (~1 KB for the base tree's 17-tuples) never exists as a whole: it is streamed
// eight bytes at a time through a 64-bit register into a small per-thread ring of 128-byte blocks in shared memory
// (word-major across the CTA), and blocks are compressed as the warp completes them (see row_leaf_kernel).
#include <string.h>

#include <vector>

#include "leaf.cuh"

namespace {

constexpr int ROW_THREADS = 256;
constexpr int ROW_RB = 2;               // blocks of 128 bytes per thread in the shared-memory ring
constexpr int ROW_RW = ROW_RB * 16;     // ... in 64-bit words
constexpr u32 MAX_ROW_PLANES = 256;
constexpr u32 MAX_ROW_TPL = 16384;

struct RowParams {
    const u8 *blob;  // device: [n_planes] plane pointers | [n_slots + 2] u32 segment offsets | modes | template bytes
    u32 n_planes, n_slots, tpl_len;
    const u8 *salts;  // n * salt_len bytes, row-major (NULL: unsalted rows)
    u32 salt_len, salt_pre_len, salt_suf_len;
    u8 salt_pre[24], salt_suf[8];
    const u32 *rows;  // row list (NULL: all rows)
    u64 n_rows, leaf_base;
    u8 *nodes;
    u32 *exc;  // exc[0] = count, exc[1 + i] = row index
};

// one BLAKE2b compression of ring block `b` of the calling thread (word w at ring[(16 b + w) * ROW_THREADS]); `h`
// lives in local memory (its address escapes), the emitter's cursor stays in registers
__device__ __noinline__ void row_compress(u64 *h, const u64 *ring, u32 b, u32 t, bool last) {
    u64 m[16], hh[8];
#pragma unroll
    for (int w = 0; w < 16; ++w) m[w] = ring[(16 * b + w) * ROW_THREADS];
#pragma unroll
    for (int i = 0; i < 8; ++i) hh[i] = h[i];
    b2b_compress(hh, m, t, last);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = hh[i];
}

// eight template bytes from an arbitrary shared-memory address (reads up to 11 bytes past it: the blob has slack)
__device__ __forceinline__ u64 read8(const u8 *p) {
    const u32 al = smem_addr(p) & 3;
    const u32 *w = reinterpret_cast<const u32 *>(p - al);
    const u32 w0 = w[0], w1 = w[1], w2 = w[2];
    return (u64)__funnelshift_r(w0, w1, 8 * al) | ((u64)__funnelshift_r(w1, w2, 8 * al) << 32);
}

// The preimage is produced in LOCKSTEP: all threads of a warp emit the same item (a template segment, then the
// integer behind it) at the same time, eight bytes per step, into their own ring of ROW_RB blocks; a block is
// compressed when EVERY thread of the warp has completed it, so the compression never runs with a partial warp
// (the first version compressed whenever a thread's own block filled up: positions differ by the integers' lengths,
// the warps diverged at every block boundary, 18 ms for the two trees of a 2^20-domain proof).  A thread may run
// ahead of the slowest one by ROW_RB * 128 - 160 = 96 bytes; rows that differ more (only contrived ones do: nine bytes
// per integer at most) leave the fast path and are redone byte by byte at the end.
__global__ void __launch_bounds__(ROW_THREADS) row_leaf_kernel(const __grid_constant__ RowParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *ring = reinterpret_cast<u64 *>(smem_raw) + threadIdx.x;  // word i of this thread at ring[i * ROW_THREADS]
    u8 *sblob = smem_raw + ROW_RW * ROW_THREADS * 8;
    const u32 blob_len = P.n_planes * 8 + (P.n_slots + 2) * 4 + P.n_planes + P.tpl_len;
    for (u32 i = threadIdx.x; i < (blob_len + 3) / 4 + 3; i += blockDim.x)
        reinterpret_cast<u32 *>(sblob)[i] = reinterpret_cast<const u32 *>(P.blob)[i];
    __syncthreads();
    const u64 *const *planes = reinterpret_cast<const u64 *const *>(sblob);
    const u32 *seg = reinterpret_cast<const u32 *>(sblob + P.n_planes * 8);
    const u8 *modes = sblob + P.n_planes * 8 + (P.n_slots + 2) * 4;
    const u8 *tpl = modes + P.n_planes;
    constexpr unsigned FULL = 0xFFFFFFFFu;

    const u64 t = (u64)blockIdx.x * ROW_THREADS + threadIdx.x;
    bool live = t < P.n_rows;  // threads without a row walk along (warp votes below) and emit nothing
    const u64 r = live ? (P.rows ? P.rows[t] : t) : 0;

    // pass 1: shape check and the total length (the frame header needs it before the first integer)
    u32 ints = 0;
    bool ok = true;
    for (u32 p = 0; p < P.n_planes; ++p) {
        const u64 v = planes[p][r];
        const u32 mode = modes[p];
        if (mode == 2) {
            ok &= v == 0;
        } else {
            ok &= mode == 0 || v != 0;
            ints += pickle_int_len(v);
        }
    }
    if (live && !ok) {
        const u32 k = atomicAdd(P.exc, 1u);
        P.exc[1 + k] = (u32)r;
        live = false;
    }
    const u32 row_len = 11 + P.tpl_len + ints;
    const u32 total = row_len + (P.salts ? P.salt_pre_len + P.salt_len + P.salt_suf_len : 0);
    u64 h[8];
    b2b_init(h);
    u64 acc = 0;
    u32 pos = 0, done = 0;  // bytes emitted by this thread; blocks compressed (the same for the whole warp)
    bool ovf = false;       // ran too far ahead of the warp: redone at the end
    // append the low `n` bytes of w (0 <= n <= 8, higher bytes of w zero) to this thread's stream
    auto append = [&](u64 w, u32 n) {
        if (live && !ovf) {
            if (pos + n > (done + ROW_RB) * 128) {
                ovf = true;
            } else {
                const u32 fill = pos & 7;
                acc |= w << (8 * fill);
                if (fill + n >= 8) {
                    ring[((pos >> 3) & (ROW_RW - 1)) * ROW_THREADS] = acc;
                    acc = fill ? w >> (64 - 8 * fill) : 0;
                }
                pos += n;
            }
        }
    };
    // compress the blocks the WHOLE warp has completed.  Warp votes: only ever called from warp-uniform control flow.
    auto flush = [&]() {
        while (__any_sync(FULL, live && !ovf) && __all_sync(FULL, !live || ovf || pos >= (done + 1) * 128)) {
            if (live && !ovf) row_compress(h, ring, done & (ROW_RB - 1), (done + 1) * 128, total == (done + 1) * 128);
            ++done;
        }
    };
    auto put = [&](u64 w, u32 n) {
        append(w, n);
        flush();
    };
    // template bytes [a, b): the same for every thread.  Whole words keep a thread's byte offset within the open word,
    // so the two shifts are set up once per segment; the warp votes every fourth word (32 bytes of the ring's slack).
    auto put_segment = [&](u32 a, u32 b) {
        const u32 sh = 8 * (pos & 7);
        u32 c = 0;
        for (; a + 8 <= b; a += 8, ++c) {
            const u64 w = read8(tpl + a);
            if (live && !ovf) {
                if (pos + 8 > (done + ROW_RB) * 128) {
                    ovf = true;
                } else {
                    ring[((pos >> 3) & (ROW_RW - 1)) * ROW_THREADS] = acc | (w << sh);
                    acc = (w >> 1) >> (63 - sh);  // w >> (64 - sh), and 0 when sh == 0
                    pos += 8;
                }
            }
            if ((c & 3) == 3) flush();
        }
        flush();
        if (a < b) put(read8(tpl + a) & (~0ull >> (64 - 8 * (b - a))), b - a);
    };
    auto put_int = [&](u64 v) {  // CPython save_long, protocol 4 (see leaf.cuh): one or two words, chosen per thread
        u64 w0, w1 = 0;
        u32 n0, n1 = 0;
        if (v < 256) {
            w0 = 0x4b | v << 8, n0 = 2;
        } else if (v < 65536) {
            w0 = 0x4d | v << 8, n0 = 3;
        } else if (v < 0x80000000ULL) {
            w0 = 0x4a | v << 8, n0 = 5;
        } else {
            const u32 nb = pickle_int_len(v) - 2;  // 5 .. 9 payload bytes, the ninth a zero
            const u32 first = nb < 6 ? nb : 6;
            n0 = 2 + first;
            w0 = (0x8a | (u64)nb << 8 | v << 16) & (~0ull >> (64 - 8 * n0));
            if (nb > 6) w1 = v >> 48, n1 = nb - 6;
        }
        put(w0, n0);
        put(w1, n1);
    };

    put(0x80 | 0x04 << 8 | 0x95 << 16, 3);
    put((u64)(row_len - 11), 8);
    u32 slot = 0;
    for (u32 p = 0; p < P.n_planes; ++p) {
        if (modes[p] == 2) continue;
        put_segment(seg[slot], seg[slot + 1]);
        put_int(planes[p][r]);
        ++slot;
    }
    put_segment(seg[slot], seg[slot + 1]);
    if (P.salts) {
        auto put_bytes = [&](const u8 *b, u32 n) {  // kernel-parameter bytes: the same for every thread
            for (u32 i = 0; i < n; i += 8) {
                u64 w = 0;
                const u32 k = n - i < 8 ? n - i : 8;
                for (u32 j = 0; j < k; ++j) w |= (u64)b[i + j] << (8 * j);
                put(w, k);
            }
        };
        put_bytes(P.salt_pre, P.salt_pre_len);
        const u8 *sp = P.salts + r * P.salt_len;
        if ((P.salt_len & 7) == 0 && ((uintptr_t)P.salts & 7) == 0) {
            for (u32 i = 0; i < P.salt_len; i += 8) put(*reinterpret_cast<const u64 *>(sp + i), 8);
        } else {
            for (u32 i = 0; i < P.salt_len; ++i) put(sp[i], 1);
        }
        put_bytes(P.salt_suf, P.salt_suf_len);
    }
    // the last, partial block: flush the open word, zero the rest, compress with the warp
    const bool fast = live && !ovf;
    if (fast && (pos & 127)) {
        u32 w = pos >> 3;
        if (pos & 7) ring[(w++ & (ROW_RW - 1)) * ROW_THREADS] = acc;
        for (; w & 15; ++w) ring[(w & (ROW_RW - 1)) * ROW_THREADS] = 0;
    }
    while (__any_sync(FULL, fast && done * 128 < total)) {
        if (fast && done * 128 < total) {
            const bool last = (done + 1) * 128 >= total;
            row_compress(h, ring, done & (ROW_RB - 1), last ? total : (done + 1) * 128, last);
        }
        ++done;
    }
    if (live && ovf) {
        // the odd row out: byte by byte through block 0 of the ring, compressing on its own
        b2b_init(h);
        acc = 0;
        pos = 0;
        auto emit = [&](u32 byte) {
            acc |= (u64)byte << (8 * (pos & 7));
            ++pos;
            if ((pos & 7) == 0) {
                ring[(((pos - 1) & 127) >> 3) * ROW_THREADS] = acc;
                acc = 0;
                if ((pos & 127) == 0) row_compress(h, ring, 0, pos, pos == total);
            }
        };
        auto emit_word = [&](u64 w, u32 n) {
            for (u32 i = 0; i < n; ++i) emit((u32)(w >> (8 * i)) & 255);
        };
        emit_word(0x80 | 0x04 << 8 | 0x95 << 16, 3);
        emit_word((u64)(row_len - 11), 8);
        u32 sl = 0;
        for (u32 p = 0; p < P.n_planes; ++p) {
            if (modes[p] == 2) continue;
            for (u32 a = seg[sl]; a < seg[sl + 1]; ++a) emit(tpl[a]);
            const u64 v = planes[p][r];
            if (v < 256) {
                emit_word(0x4b | v << 8, 2);
            } else if (v < 65536) {
                emit_word(0x4d | v << 8, 3);
            } else if (v < 0x80000000ULL) {
                emit_word(0x4a | v << 8, 5);
            } else {
                const u32 nb = pickle_int_len(v) - 2;
                emit(0x8a);
                emit(nb);
                for (u32 i = 0; i < nb; ++i) emit(i < 8 ? (u32)(v >> (8 * i)) & 255 : 0);
            }
            ++sl;
        }
        for (u32 a = seg[sl]; a < seg[sl + 1]; ++a) emit(tpl[a]);
        if (P.salts) {
            for (u32 i = 0; i < P.salt_pre_len; ++i) emit(P.salt_pre[i]);
            const u8 *sp = P.salts + r * P.salt_len;
            for (u32 i = 0; i < P.salt_len; ++i) emit(sp[i]);
            for (u32 i = 0; i < P.salt_suf_len; ++i) emit(P.salt_suf[i]);
        }
        if (pos & 127) {
            u32 w = (pos & 127) >> 3;
            if (pos & 7) ring[w++ * ROW_THREADS] = acc;
            for (; w < 16; ++w) ring[w * ROW_THREADS] = 0;
            row_compress(h, ring, 0, total, true);
        }
    }
    if (live) store_digest(P.nodes, P.leaf_base + r, h);
}

}  // namespace

extern "C" int b2s_merkle_rows(const uint64_t *const *h_planes, const uint8_t *h_modes, uint32_t n_planes, uint64_t n,
                               const uint8_t *h_tpl, const uint32_t *h_seg_off, uint32_t n_slots,
                               const uint8_t *d_salts, uint32_t salt_len, const uint8_t *h_salt_prefix,
                               uint32_t salt_prefix_len, const uint8_t *h_salt_suffix, uint32_t salt_suffix_len,
                               const uint32_t *d_rows, uint64_t n_rows, uint8_t *d_nodes, int build_upper,
                               uint32_t *h_exceptions, uint32_t *h_n_exceptions, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (h_n_exceptions) *h_n_exceptions = 0;
    if (n == 0 || (n & (n - 1)) || n > 0xFFFFFFFFull) {
        b2s_set_error("merkle_rows: leaf count must be a power of two below 2^32, got %llu", (unsigned long long)n);
        return B2S_ERR_ARG;
    }
    if (n_planes == 0 || n_planes > MAX_ROW_PLANES || n_slots > n_planes) {
        b2s_set_error("merkle_rows: %u planes / %u integer slots", n_planes, n_slots);
        return B2S_ERR_ARG;
    }
    u32 emitted = 0;
    for (u32 p = 0; p < n_planes; ++p) {
        if (h_modes[p] > 2) {
            b2s_set_error("merkle_rows: plane mode must be 0 (emit), 1 (emit, non-zero) or 2 (absent, zero)");
            return B2S_ERR_ARG;
        }
        emitted += h_modes[p] != 2;
    }
    const u32 tpl_len = h_seg_off[n_slots + 1];
    if (emitted != n_slots || h_seg_off[0] != 0 || tpl_len > MAX_ROW_TPL) {
        b2s_set_error("merkle_rows: template has %u slots for %u emitted planes (%u template bytes)", n_slots, emitted,
                      tpl_len);
        return B2S_ERR_ARG;
    }
    for (u32 j = 0; j <= n_slots; ++j)
        if (h_seg_off[j + 1] < h_seg_off[j]) {
            b2s_set_error("merkle_rows: bad segment offsets");
            return B2S_ERR_ARG;
        }
    if (d_salts && (salt_prefix_len > 24 || salt_suffix_len > 8)) {
        b2s_set_error("merkle_rows: salt pickle frame of %u + %u bytes", salt_prefix_len, salt_suffix_len);
        return B2S_ERR_ARG;
    }
    const u64 count = d_rows ? n_rows : n;
    if (count == 0) return 0;

    std::vector<u8> blob((size_t)n_planes * 8 + (n_slots + 2) * 4 + n_planes + tpl_len + 32, 0);  // slack: read8 / word copies
    memcpy(blob.data(), h_planes, (size_t)n_planes * 8);
    memcpy(blob.data() + n_planes * 8, h_seg_off, (n_slots + 2) * 4);
    memcpy(blob.data() + n_planes * 8 + (n_slots + 2) * 4, h_modes, n_planes);
    memcpy(blob.data() + n_planes * 8 + (n_slots + 2) * 4 + n_planes, h_tpl, tpl_len);
    u8 *d_blob = nullptr;
    u32 *d_exc = nullptr;
    B2S_CUDA(cudaMallocAsync(&d_blob, blob.size(), st));
    B2S_CUDA(cudaMallocAsync(&d_exc, (count + 1) * sizeof(u32), st));
    B2S_CUDA(cudaMemcpyAsync(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
    B2S_CUDA(cudaMemsetAsync(d_exc, 0, sizeof(u32), st));

    RowParams P;
    memset(&P, 0, sizeof(P));
    P.blob = d_blob;
    P.n_planes = n_planes;
    P.n_slots = n_slots;
    P.tpl_len = tpl_len;
    P.salts = d_salts;
    P.salt_len = salt_len;
    P.salt_pre_len = d_salts ? salt_prefix_len : 0;
    P.salt_suf_len = d_salts ? salt_suffix_len : 0;
    if (d_salts) {
        memcpy(P.salt_pre, h_salt_prefix, salt_prefix_len);
        memcpy(P.salt_suf, h_salt_suffix, salt_suffix_len);
    }
    P.rows = d_rows;
    P.n_rows = count;
    P.leaf_base = n;
    P.nodes = d_nodes;
    P.exc = d_exc;
    const size_t smem = (size_t)ROW_RW * ROW_THREADS * 8 + ((blob.size() + 15) & ~(size_t)15);
    static bool attr[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 15]) {
        B2S_CUDA(cudaFuncSetAttribute(row_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      ROW_RW * ROW_THREADS * 8 + MAX_ROW_PLANES * 13 + MAX_ROW_TPL + 64));
        attr[dev & 15] = true;
    }
    if (!d_rows) B2S_CUDA(cudaMemsetAsync(d_nodes, 0, 64, st));  // slot 0 is never a node
    row_leaf_kernel<<<(unsigned)((count + ROW_THREADS - 1) / ROW_THREADS), ROW_THREADS, smem, st>>>(P);
    B2S_LAUNCHED();
    u32 n_exc = 0;
    B2S_CUDA(cudaMemcpyAsync(&n_exc, d_exc, sizeof(u32), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));  // also: the host blob may go out of scope
    if (n_exc && h_exceptions)
        B2S_CUDA(cudaMemcpyAsync(h_exceptions, d_exc + 1, (size_t)n_exc * sizeof(u32), cudaMemcpyDeviceToHost, st));
    B2S_CUDA(cudaStreamSynchronize(st));
    B2S_CUDA(cudaFreeAsync(d_blob, st));
    B2S_CUDA(cudaFreeAsync(d_exc, st));
    if (h_n_exceptions) *h_n_exceptions = n_exc;
    if (build_upper && n_exc == 0 && n > 1) return merkle_upper_run(d_nodes, n, st);
    return 0;
}
