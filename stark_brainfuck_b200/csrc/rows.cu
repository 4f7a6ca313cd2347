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
// One thread per row.  The preimage (~1 KB for the base tree's 17-tuples) is streamed: bytes are packed
// into a 64-bit register, words into the thread's 128-byte block in shared memory (word-major across the
// CTA), and every full block is compressed at once -- the message never exists as a whole.
#include <string.h>

#include <vector>

#include "leaf.cuh"

namespace {

constexpr int ROW_THREADS = 128;
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

// one BLAKE2b compression of the thread's block; `h` lives in local memory (its address escapes), the
// emitter's cursor stays in registers
__device__ __noinline__ void row_compress(u64 *h, const u64 *blk, u32 t, bool last) {
    u64 m[16], hh[8];
#pragma unroll
    for (int w = 0; w < 16; ++w) m[w] = blk[w * ROW_THREADS];
#pragma unroll
    for (int i = 0; i < 8; ++i) hh[i] = h[i];
    b2b_compress(hh, m, t, last);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = hh[i];
}

__global__ void __launch_bounds__(ROW_THREADS) row_leaf_kernel(const __grid_constant__ RowParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *blocks = reinterpret_cast<u64 *>(smem_raw);  // [16][ROW_THREADS]
    u8 *sblob = smem_raw + 16 * ROW_THREADS * 8;
    const u32 blob_len = P.n_planes * 8 + (P.n_slots + 2) * 4 + P.n_planes + P.tpl_len;
    for (u32 i = threadIdx.x; i < (blob_len + 3) / 4; i += blockDim.x)
        reinterpret_cast<u32 *>(sblob)[i] = reinterpret_cast<const u32 *>(P.blob)[i];
    __syncthreads();
    const u64 *const *planes = reinterpret_cast<const u64 *const *>(sblob);
    const u32 *seg = reinterpret_cast<const u32 *>(sblob + P.n_planes * 8);
    const u8 *modes = sblob + P.n_planes * 8 + (P.n_slots + 2) * 4;
    const u8 *tpl = modes + P.n_planes;

    const u64 t = (u64)blockIdx.x * ROW_THREADS + threadIdx.x;
    if (t >= P.n_rows) return;
    const u64 r = P.rows ? P.rows[t] : t;

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
    if (!ok) {
        const u32 k = atomicAdd(P.exc, 1u);
        P.exc[1 + k] = (u32)r;
        return;
    }
    const u32 row_len = 11 + P.tpl_len + ints;
    u64 h[8];
    b2b_init(h);
    u64 acc = 0;
    u32 pos = 0;
    const u32 total = row_len + (P.salts ? P.salt_pre_len + P.salt_len + P.salt_suf_len : 0);
    u64 *blk = blocks + threadIdx.x;  // word w of this thread's block at blk[w * ROW_THREADS]
    auto emit = [&](u32 byte) {
        acc |= (u64)byte << (8 * (pos & 7));
        ++pos;
        if ((pos & 7) == 0) {
            blk[(((pos - 1) & 127) >> 3) * ROW_THREADS] = acc;
            acc = 0;
            if ((pos & 127) == 0) row_compress(h, blk, pos, pos == total);
        }
    };
    auto emit_int = [&](u64 v) {
        if (v < 256) {
            emit(0x4b);
            emit((u32)v);
        } else if (v < 65536) {
            emit(0x4d);
            emit((u32)v & 255);
            emit((u32)(v >> 8));
        } else if (v < 0x80000000ULL) {
            emit(0x4a);
            for (int i = 0; i < 4; ++i) emit((u32)(v >> (8 * i)) & 255);
        } else {
            const u32 nb = pickle_int_len(v) - 2;
            emit(0x8a);
            emit(nb);
            for (u32 i = 0; i < nb; ++i) emit(i < 8 ? (u32)(v >> (8 * i)) & 255 : 0);
        }
    };

    emit(0x80);
    emit(0x04);
    emit(0x95);
    const u64 body = row_len - 11;
    for (int i = 0; i < 8; ++i) emit((u32)(body >> (8 * i)) & 255);
    u32 slot = 0;
    for (u32 p = 0; p < P.n_planes; ++p) {
        if (modes[p] == 2) continue;
        for (u32 a = seg[slot]; a < seg[slot + 1]; ++a) emit(tpl[a]);
        emit_int(planes[p][r]);
        ++slot;
    }
    for (u32 a = seg[slot]; a < seg[slot + 1]; ++a) emit(tpl[a]);
    if (P.salts) {
        for (u32 i = 0; i < P.salt_pre_len; ++i) emit(P.salt_pre[i]);
        const u8 *sp = P.salts + r * P.salt_len;
        for (u32 i = 0; i < P.salt_len; ++i) emit(sp[i]);
        for (u32 i = 0; i < P.salt_suf_len; ++i) emit(P.salt_suf[i]);
    }
    if (pos & 127) {  // partial final block: flush the open word, zero the rest
        u32 w = (pos & 127) >> 3;
        if (pos & 7) blk[w++ * ROW_THREADS] = acc;
        for (; w < 16; ++w) blk[w * ROW_THREADS] = 0;
        row_compress(h, blk, total, true);
    }
    store_digest(P.nodes, P.leaf_base + r, h);
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

    std::vector<u8> blob((size_t)n_planes * 8 + (n_slots + 2) * 4 + n_planes + tpl_len + 4, 0);
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
    const size_t smem = 16 * ROW_THREADS * 8 + ((blob.size() + 15) & ~(size_t)15);
    static bool attr[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 15]) {
        B2S_CUDA(cudaFuncSetAttribute(row_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      16 * ROW_THREADS * 8 + MAX_ROW_PLANES * 13 + MAX_ROW_TPL + 64));
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
