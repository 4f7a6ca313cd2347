// dist.cu -- the exchange step of the four-step NTT across GPUs.
//
// n = n1*n2, j = j1 + n1*j2, k = k1*n2 + k2.  Rank g owns q = n1/G columns j1 and holds them as
// planes P[j1_local][j2].  After the local length-n2 transforms A[j1_local][k2] every element is
// multiplied by the twiddle omega^(j1*k2) and sent to the rank that owns row k2; there the
// planes C[k2_local][j1] feed the local length-n1 transforms.  This kernel does twiddle +
// transpose + placement in one pass and writes through a table of destination pointers:
//   * NCCL exchange : all pointers lie in the local send buffer, layout [peer][k2_local][j1_local]
//   * peer stores   : pointer r is rank r's receive buffer mapped over NVLink (symmetric memory);
//                     rows land directly at C_r[k2_local][g*q + j1_local], so the transfer is part
//                     of the compute kernel and overlaps it tile by tile.
#include "common.h"
#include "glmont.cuh"

namespace {

constexpr int MAX_PEERS = 16;

struct DistParams {
    const u64 *in;   // [rows][in_stride]
    u64 in_stride;   // elements
    u32 rows, cols;  // rows = owned columns j1 (q), cols = n2
    u64 row_base;    // global index of local row 0 (g*q)
    u64 tw_mul;      // twiddle exponent = tw_mul * (row_base + r) * c
    u32 log_cols_per_peer;
    u64 out_row_stride, out_col_offset;
    u64 *out[MAX_PEERS];
    u64 w_sq[32];
};

// tile 32 (rows) x 32 (cols); thread (tx, ty) with ty < 8 owns rows ty, ty+8, ty+16, ty+24 of the
// row block and the column residue tx, and walks the column tiles with a geometric twiddle.
__global__ void __launch_bounds__(256) dist_twiddle_transpose_kernel(const __grid_constant__ DistParams P) {
    __shared__ u64 tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const u32 r0 = blockIdx.x * 32;
    u64 tw[4], ratio[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u64 j1 = P.row_base + r0 + ty + 8 * i;
        // both in Montgomery form: mont_mul(x, tw) = x * omega^e, mont_mul(tw, ratio) steps tw by 32 columns
        tw[i] = gl_to_mont(gl_pow_sq(P.w_sq, P.tw_mul * j1 * (u64)tx));
        ratio[i] = gl_to_mont(gl_pow_sq(P.w_sq, P.tw_mul * j1 * 32));
    }
    for (u32 c0 = 0; c0 < P.cols; c0 += 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 r = r0 + ty + 8 * i, c = c0 + tx;
            u64 v = 0;
            if (r < P.rows && c < P.cols) v = lcanon(mont_mul(P.in[(u64)r * P.in_stride + c], tw[i]));
            tile[ty + 8 * i][tx] = v;
            tw[i] = mont_mul(tw[i], ratio[i]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 c = c0 + ty + 8 * i, r = r0 + tx;  // transposed: tx runs along the rows now
            if (r < P.rows && c < P.cols) {
                const u32 peer = c >> P.log_cols_per_peer;
                const u32 cl = c & ((1u << P.log_cols_per_peer) - 1);
                P.out[peer][(u64)cl * P.out_row_stride + P.out_col_offset + r] = tile[tx][ty + 8 * i];
            }
        }
        __syncthreads();
    }
}

// out[b][a][:] = in[a][b][:] for runs of c contiguous elements
__global__ void __launch_bounds__(256) block_permute_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u32 A,
                                                            u32 B, u32 C) {
    const u64 total = (u64)A * B * C;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        const u32 c = (u32)(i % C);
        const u64 ab = i / C;
        const u32 a = (u32)(ab % A), b = (u32)(ab / A);  // i indexes out[b][a][c]
        out[i] = in[((u64)a * B + b) * C + c];
    }
}

}  // namespace

extern "C" int b2s_dist_twiddle_transpose(const uint64_t *d_in, uint64_t in_stride, uint32_t rows, uint32_t cols,
                                          uint64_t row_base, uint64_t omega, uint64_t tw_mul, uint64_t *const *out_ptrs,
                                          uint32_t n_peers, uint64_t out_row_stride, uint64_t out_col_offset,
                                          void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_peers == 0 || n_peers > MAX_PEERS || (n_peers & (n_peers - 1)) || cols % n_peers || rows == 0 || cols == 0) {
        b2s_set_error("dist_twiddle_transpose: bad shape (rows %u cols %u peers %u)", rows, cols, n_peers);
        return B2S_ERR_ARG;
    }
    const u32 cpp = cols / n_peers;
    if (cpp & (cpp - 1)) {
        b2s_set_error("dist_twiddle_transpose: columns per peer must be a power of two");
        return B2S_ERR_ARG;
    }
    DistParams P;
    P.in = d_in;
    P.in_stride = in_stride;
    P.rows = rows;
    P.cols = cols;
    P.row_base = row_base;
    P.tw_mul = tw_mul;
    P.log_cols_per_peer = ilog2_u64(cpp);
    P.out_row_stride = out_row_stride;
    P.out_col_offset = out_col_offset;
    for (u32 i = 0; i < MAX_PEERS; ++i) P.out[i] = i < n_peers ? out_ptrs[i] : nullptr;
    u64 sq = omega % GL_P;
    for (int b = 0; b < 32; ++b) {
        P.w_sq[b] = sq;
        sq = gl_mul(sq, sq);
    }
    dist_twiddle_transpose_kernel<<<dim3((rows + 31) / 32), dim3(32, 8), 0, st>>>(P);
    B2S_LAUNCHED();
    return 0;
}

extern "C" int b2s_block_permute(const uint64_t *d_in, uint64_t *d_out, uint32_t A, uint32_t B, uint32_t C,
                                 void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const u64 total = (u64)A * B * C;
    if (total == 0) return 0;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    block_permute_kernel<<<blocks, 256, 0, st>>>(d_in, d_out, A, B, C);
    B2S_LAUNCHED();
    return 0;
}
