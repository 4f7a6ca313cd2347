/*
 * b2s.h -- C ABI of libb2s.so, the sm_100a polynomial / FRI / Merkle engine that sits
 * behind the call surface of aszepieniec/stark-brainfuck's hot path.
 *
 * The reference has no FFI layer (it is pure Python); each entry point below names the
 * reference function it replaces (paths relative to the reference checkout).  The
 * reference-side binding is a ctypes stub, shown in INTEGRATION.md and implemented in
 * stark_brainfuck_b200/_lib.py.
 *
 * Conventions
 *  - Field elements are canonical uint64 values in [0, p), p = 2^64 - 2^32 + 1
 *    (code/algebra.py:110-115).  Montgomery or lazy forms never cross this boundary.
 *  - An extension-field vector (code/extension_field.py, F_p[X]/(X^3 - X + 1)) is three
 *    planes c0, c1, c2 of n uint64 each; trimmed (absent) coefficients are zero.
 *  - `d_` pointers are DEVICE pointers owned by the caller (in the Python glue: torch
 *    uint8/int64 tensors used as byte buffers).  `h_` pointers are HOST pointers.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device-pointer
 *    entry points only enqueue work; they do not synchronise.  `_host` entry points copy
 *    in, run, copy out and synchronise before returning.
 *  - Every function returns 0 on success or a negative B2S_ERR_* code and never throws;
 *    b2s_last_error() returns a thread-local message for the last failure.
 *  - Precondition failures that the reference reports with `assert` come back as
 *    B2S_ERR_ASSERT_*; the glue turns them into AssertionError.
 */
#ifndef B2S_H
#define B2S_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_ERR_CUDA (-1)
#define B2S_ERR_ARG (-2)
#define B2S_ERR_NOMEM (-3)
#define B2S_ERR_NCCL (-4)
#define B2S_ERR_ASSERT_NPO2 (-10)      /* code/ntt.py:5-6   "cannot compute ntt of non-power-of-two sequence" */
#define B2S_ERR_ASSERT_ROOT (-11)      /* code/ntt.py:13-14 "primitive root must be nth root of unity" */
#define B2S_ERR_ASSERT_PRIMITIVE (-12) /* code/ntt.py:15-16 "primitive root is not primitive nth root of unity" */

#define B2S_TPL_MAX_BYTES 2048

/* Byte templates of pickle.dumps(leaf) for field-element Merkle leaves
 * (code/merkle.py:29-32 hashes pickle.dumps(leaf)).  The glue derives them at start-up
 * by pickling marker elements of the caller's own classes, so module names, memo
 * indices and the shared-`field` identity pattern are whatever the caller's objects
 * produce.  Template k is used for an element with k coefficients; its k+1 byte
 * segments are interleaved with the k pickled integers:
 *     PROTO 4 | FRAME(len) | seg0 INT(c0) seg1 INT(c1) ... seg_k          (seg_k ends in STOP)
 */
typedef struct b2s_leaf_templates {
    uint32_t n_slots;       /* 1: BaseFieldElement leaves; 3: ExtensionFieldElement leaves */
    uint32_t trim;          /* 1: k = number of coefficients left after trimming trailing zeros
                               (code/extension_field.py:6-9); 0: k = n_slots always */
    uint32_t seg_off[4][5]; /* template k, segment j = bytes[seg_off[k][j] .. seg_off[k][j+1]) */
    uint8_t bytes[B2S_TPL_MAX_BYTES];
} b2s_leaf_templates;

/* ---- library / device management -------------------------------------------------- */
int b2s_version(void);
const char *b2s_last_error(void);
/* Select the CUDA device for the calling thread and create the per-device caches.
 * Fails (B2S_ERR_CUDA) when no sm_100 device is visible: there is no CPU fallback.
 * Scratch memory comes from the device's stream-ordered pool, which keeps up to 2 GiB of freed scratch
 * (environment B2S_POOL_THRESHOLD_MB overrides the bound; b2s_trim() returns everything to the driver). */
int b2s_init(int device);
int b2s_shutdown(void);
/* Return cached device memory to the driver: the stream-ordered scratch pool (which otherwise keeps up to 2 GiB
 * of freed scratch for reuse) and the twiddle-table cache.  Synchronises the device.  For callers that share the
 * device with another allocator and run into its out-of-memory condition. */
int b2s_trim(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t b2s_launch_count(void);
int b2s_device_sm_count(void);

/* ---- scalar helpers (host side, exact) -------------------------------------------- */
/* code/algebra.py:89-108, :39-46 */
uint64_t b2s_gl_mul(uint64_t a, uint64_t b);
uint64_t b2s_gl_pow(uint64_t a, uint64_t e);
uint64_t b2s_gl_inv(uint64_t a);

/* ---- NTT family --------------------------------------------------------------------
 * One entry point covers code/ntt.py:4-23 (ntt), :26-42 (intt), :164-168
 * (fast_coset_evaluate), :171-174 (fast_coset_interpolate), code/fri.py:26-44
 * (Fri.Domain.evaluate/xevaluate/interpolate/xinterpolate) and the scale step of
 * code/univariate.py:168-169 that those wrappers fuse.
 *
 *   inverse == 0:  out[k] = sum_{j < n_in} (offset^j * in[j]) * omega^(j*k),  k < n = 2^log_n
 *                  (coefficients beyond n_in are the zero padding of code/fri.py:28-29)
 *   inverse != 0:  out[j] = offset^(-j) * n^(-1) * sum_k in[k] * omega^(-j*k)     (n_in == n)
 *
 * offset == 1 gives the plain ntt / intt.  Natural order in and out.  `n_planes`
 * independent vectors (1 for a base-field vector, 3 for an extension-field vector, any
 * number for a batch of columns) are transformed in one call; plane q starts at
 * d_in + q*in_stride and d_out + q*out_stride (strides in elements).  d_in == d_out is
 * allowed when in_stride == out_stride.  Scratch comes from the stream-ordered pool.
 * Checks the asserts of code/ntt.py:13-16 on omega (B2S_ERR_ASSERT_*). */
int b2s_ntt(const uint64_t *d_in, uint64_t in_stride, uint32_t n_in, uint64_t *d_out, uint64_t out_stride,
            uint32_t log_n, uint32_t n_planes, uint64_t omega, uint64_t offset, int inverse, void *stream);
/* Same through host buffers: H2D, transform, D2H, synchronise (the e2e path). */
int b2s_ntt_host(const uint64_t *h_in, uint64_t in_stride, uint32_t n_in, uint64_t *h_out, uint64_t out_stride,
                 uint32_t log_n, uint32_t n_planes, uint64_t omega, uint64_t offset, int inverse);

/* code/univariate.py:168-169 Polynomial.scale: out[i] = factor^i * in[i].
 * n_planes == 1: base-field coefficients, factor[0] used.  n_planes == 3: extension-field
 * coefficients with an extension-field factor (factor[0..2]). */
int b2s_scale(const uint64_t *d_in, uint64_t in_stride, uint64_t *d_out, uint64_t out_stride, uint64_t n,
              uint32_t n_planes, const uint64_t factor[3], void *stream);

/* code/univariate.py:145-154 Polynomial.evaluate_domain on ARBITRARY points (running-power
 * evaluation per point).  coeff_planes / point_planes are 1 or 3; the output has
 * max(coeff_planes, point_planes) planes.  (Points that form a coset offset*omega^k are
 * routed to b2s_ntt by the glue instead.) */
int b2s_eval_points(const uint64_t *d_coeffs, uint64_t coeff_stride, uint32_t coeff_planes, uint64_t n_coeffs,
                    const uint64_t *d_points, uint64_t point_stride, uint32_t point_planes, uint64_t n_points,
                    uint64_t *d_out, uint64_t out_stride, void *stream);

/* ---- Merkle ------------------------------------------------------------------------
 * code/merkle.py:8-41: BLAKE2b-512 tree, heap layout.  d_nodes holds 2*n slots of 64
 * bytes: slot n+i = blake2b(pickle.dumps(leaf_i)), slot k = blake2b(slot 2k | slot 2k+1),
 * slot 1 = root, slot 0 unused (zeroed).  Leaves are field elements given as planes
 * (n_slots planes of n values, n a power of two); their pickle preimages are emitted on
 * the device from `tpl`. */
int b2s_merkle_field(const uint64_t *d_planes, uint64_t plane_stride, uint64_t n, const b2s_leaf_templates *tpl,
                     uint8_t *d_nodes, void *stream);
/* Arbitrary picklable leaves (code/test_merkle.py:57-61): the caller pickles on the host;
 * leaf i is d_bytes[d_offsets[i] .. d_offsets[i+1]).  n_leafs need not be a power of two:
 * unused leaf slots behave as the reference's 32-byte zero placeholders
 * (code/merkle.py:26).  d_nodes holds 2*npo2 slots. */
int b2s_merkle_blobs(const uint8_t *d_bytes, const uint64_t *d_offsets, uint64_t n_leafs, uint64_t npo2,
                     uint8_t *d_nodes, void *stream);
/* Leaves that are ROWS of several codewords -- the zipped, salted leaves of BrainfuckStark.prove
 * (code/brainfuck_stark.py:178-180, :197-199; code/salted_merkle.py:25-35):
 *     slot n+r = blake2b(pickle.dumps(row_r) | pickle.dumps(salt_r)),   row_r = (cw_0[r], cw_1[r], ...).
 * The glue derives ONE byte template per tree from a sample row of the caller's own objects (the elements of
 * a row carry different `field` objects, so memo indices are per tree); the device splices the integers:
 *     PROTO 4 | FRAME(len) | seg_0 INT(v_0) seg_1 ... INT(v_{S-1}) seg_S | salt_prefix salt_r salt_suffix
 * seg_j = h_tpl[h_seg_off[j] .. h_seg_off[j+1]) (n_slots + 2 offsets).  h_planes: HOST array of n_planes DEVICE
 * pointers (n values each), in row order; h_modes[p]: 0 = integer slot, 1 = integer slot that must be non-zero
 * (top coefficient of an extension-field element), 2 = no slot, value must be zero (a trimmed coefficient,
 * code/extension_field.py:6-9).  A row that violates its modes has a different pickle shape: it is NOT hashed
 * and its index is returned in h_exceptions (capacity = number of rows processed); the caller hashes those rows
 * with the template of their shape through d_rows (DEVICE list of n_rows row indices; NULL = all n rows).
 * d_salts: n * salt_len bytes (NULL: unsalted rows).  build_upper != 0 and no exceptions: the inner nodes
 * (code/salted_merkle.py:38-44) are built in the same call.  n must be a power of two.  Synchronises. */
int b2s_merkle_rows(const uint64_t *const *h_planes, const uint8_t *h_modes, uint32_t n_planes, uint64_t n,
                    const uint8_t *h_tpl, const uint32_t *h_seg_off, uint32_t n_slots, const uint8_t *d_salts,
                    uint32_t salt_len, const uint8_t *h_salt_prefix, uint32_t salt_prefix_len,
                    const uint8_t *h_salt_suffix, uint32_t salt_suffix_len, const uint32_t *d_rows, uint64_t n_rows,
                    uint8_t *d_nodes, int build_upper, uint32_t *h_exceptions, uint32_t *h_n_exceptions, void *stream);
/* code/merkle.py:35-41 alone: the inner nodes above `npo2` digests that the caller has placed in
 * slots [npo2, 2*npo2) of d_nodes (multi-GPU trees: the top levels over the subtree roots that the
 * ranks exchanged). */
int b2s_merkle_upper(uint8_t *d_nodes, uint64_t npo2, void *stream);
/* code/merkle.py:46-52 open(): copies the `depth` sibling digests of each index, leaf
 * level first, to host memory: h_paths[q*depth*64 ...].  Synchronises. */
int b2s_merkle_open(const uint8_t *d_nodes, uint64_t npo2, const uint64_t *h_indices, uint32_t n_indices,
                    uint8_t *h_paths, void *stream);

/* ---- FRI ---------------------------------------------------------------------------
 * One commit round of code/fri.py:91-139 on an extension-field codeword of N values:
 *   d_next[i] = 2^-1 * ((1 + alpha/(offset*omega^i)) * cw[i] + (1 - alpha/(offset*omega^i)) * cw[N/2+i])
 * for i < N/2 (code/fri.py:127-128).  If d_next_nodes != NULL the Merkle tree of the
 * folded codeword (code/fri.py:108 of the NEXT round) is built in the same call
 * (leaf hashing fused with the fold). */
int b2s_fri_fold(const uint64_t *d_cw, uint64_t cw_stride, uint64_t N, const uint64_t alpha[3], uint64_t offset,
                 uint64_t omega, uint64_t *d_next, uint64_t next_stride, const b2s_leaf_templates *tpl,
                 uint8_t *d_next_nodes, void *stream);
/* Gather elements of planes at the given indices to the host (code/fri.py:150, :169 read
 * tree.leafs[i] / last_codeword[i]): h_out[q*n_planes + plane].  Synchronises. */
int b2s_gather(const uint64_t *d_planes, uint64_t plane_stride, uint32_t n_planes, const uint64_t *h_indices,
               uint32_t n_indices, uint64_t *h_out, void *stream);

/* ---- quotient codewords ---------------------------------------
 * code/table.py:155-178 boundary_quotients, :190-236 transition_quotients, :253-286
 * terminal_quotients and code/permutation_argument.py:11-20 quotient: for every point
 * x_i = offset * omega^i of the FRI domain (code/fri.py:20-21) and every constraint c
 *     out[c][i] = ( sum_m coeff[m] * prod_f var(v_f)[i] ^ e_f ) * zinv(x_i)
 * which is code/multivariate.py:105-116 MPolynomial.evaluate at the point
 *     var(v)[i] = cw[v][i]                        for v <  width
 *               = cw[v - width][(i + shift) % N]  for v >= width   (the "next row" variables)
 * times the inverse zerofier:
 *     B2S_ZEROFIER_BOUNDARY    (x_i - 1)^-1                               code/table.py:161-163
 *     B2S_ZEROFIER_TRANSITION  (x_i^height - 1)^-1 * (x_i - omicron_inv)  code/table.py:194-201
 *                              (height == 0: zero, as in the reference)
 *     B2S_ZEROFIER_TERMINAL    (x_i - omicron_inv)^-1                     code/table.py:256-259
 * d_cw: `width` extension-field codewords, codeword v = planes d_cw + (3 v + s) * N, s < 3.
 * The constraint program is given in HOST memory: constraint c owns the monomials
 * h_mono_off[c] .. h_mono_off[c+1]; monomial m has coefficient h_coeffs[3 m .. 3 m + 2] and
 * factors h_factors[m * max_factors + f] = (variable << 8) | exponent, exponent 0 = unused.
 * d_out: n_constraints codewords in the same plane layout.  *h_zero_flag is set to 1 when a
 * zerofier vanishes on the domain (the reference's batch_inverse asserts, code/ntt.py:178-179).
 * h_zero_flag may be NULL when the caller has ruled that out itself (offset^N != 1 puts every point
 * outside the subgroup the zerofiers' roots live in): the call then reads nothing back and does not
 * synchronise, so consecutive tables' kernels run back to back.
 * h_base_columns (HOST, `width` bytes, or NULL): non-zero for a codeword the caller KNOWS to be a lifted
 * base-field column, i.e. with all-zero planes 1 and 2 (every Table.extend of the reference lifts its base
 * codewords, e.g. code/io_table.py:106-107); their factors are multiplied in the base field (1 multiplication
 * instead of 9).  NULL: the library scans the columns itself (one read-back).  Synchronises unless
 * h_zero_flag is NULL and h_base_columns is given. */
#define B2S_ZEROFIER_BOUNDARY 1
#define B2S_ZEROFIER_TRANSITION 2
#define B2S_ZEROFIER_TERMINAL 3
int b2s_quotients(const uint64_t *d_cw, uint64_t N, uint32_t width, uint64_t shift, uint32_t n_constraints,
                  const uint32_t *h_mono_off, const uint64_t *h_coeffs, const uint32_t *h_factors, uint32_t max_factors,
                  uint32_t zerofier_kind, uint64_t height, uint64_t omicron_inv, uint64_t offset, uint64_t omega,
                  uint64_t *d_out, int *h_zero_flag, const uint8_t *h_base_columns, void *stream);

/* code/fri.py:141-176, the query phase: the codeword elements (`tree.leafs[i]`, :150/:169) and authentication
 * paths (`tree.open(i)`, code/merkle.py:46-52) of SEVERAL trees in one launch and one synchronisation.  Set s
 * has h_counts[s] indices (all sets concatenated in h_indices); h_planes[s] / h_nodes[s] are DEVICE addresses,
 * either may be NULL (then the set contributes zeros to h_values / nothing to h_paths).  h_values: n_planes
 * values per index; h_paths: log2(h_npo2[s]) * 64 bytes per index of a set with a tree, in index order.
 * All h_* arrays live in HOST memory. */
int b2s_open_multi(const uint64_t *const *h_planes, const uint64_t *h_plane_strides, uint32_t n_planes,
                   const uint8_t *const *h_nodes, const uint64_t *h_npo2, const uint32_t *h_counts,
                   const uint64_t *h_indices, uint32_t n_sets, uint64_t *h_values, uint8_t *h_paths, void *stream);

/* ---- nonlinear combination codeword --------------------------------
 * code/brainfuck_stark.py:241-298: the reference builds, for every base / extension / quotient
 * codeword c, the terms c and x^shift * c, and sums weight * term over all of them element by
 * element.  Here
 *     out[j] = sum_c (wa_c + wb_c * x_j^shift_c) * col_c[j],   x_j = offset * omega^j,  j < N.
 * Column c is a base-field plane (h_planes[c] == 1; lifted like ExtensionField.lift,
 * code/extension_field.py:113-116) or an extension-field plane triple (h_planes[c] == 3, planes
 * h_strides[c] elements apart) anywhere in device memory: h_cols[c] is its DEVICE address; all
 * h_* arrays live in HOST memory.  h_wa / h_wb: 3 coefficients per column; a column whose wb is
 * zero has no shifted term (the randomizer codeword, :242).  d_out: 3 planes, out_stride apart.
 * Synchronises. */
int b2s_combination(const uint64_t *const *h_cols, const uint64_t *h_strides, const uint32_t *h_planes,
                    const uint64_t *h_wa, const uint64_t *h_wb, const uint64_t *h_shifts, uint32_t n_cols, uint64_t N,
                    uint64_t offset, uint64_t omega, uint64_t *d_out, uint64_t out_stride, void *stream);

/* ---- multi-GPU exchange step of the four-step NTT (no counterpart in the single-process
 * reference) ------------------------------------------------------------
 * n = n1*n2, j = j1 + n1*j2.  The caller owns `rows` = n1/G columns j1 (first one: row_base) as
 * planes d_in[j1_local][k2] (after the local length-n2 transforms, cols = n2).  Writes
 *     out_ptrs[k2 / (cols/n_peers)][(k2 % (cols/n_peers)) * out_row_stride + out_col_offset + j1_local]
 *         = in[j1_local][k2] * omega^(tw_mul * j1 * k2)
 * i.e. twiddle + transpose + placement in one kernel.  out_ptrs are device pointers: slices of a
 * local send buffer (NCCL all-to-all follows) or peer buffers mapped over NVLink (the kernel
 * then IS the exchange).  out_ptrs itself is a HOST array of n_peers <= 16 pointers. */
int b2s_dist_twiddle_transpose(const uint64_t *d_in, uint64_t in_stride, uint32_t rows, uint32_t cols,
                               uint64_t row_base, uint64_t omega, uint64_t tw_mul, uint64_t *const *out_ptrs,
                               uint32_t n_peers, uint64_t out_row_stride, uint64_t out_col_offset, void *stream);
/* d_out[b][a][0..C) = d_in[a][b][0..C): reorders the blocks received by an all-to-all. */
int b2s_block_permute(const uint64_t *d_in, uint64_t *d_out, uint32_t A, uint32_t B, uint32_t C, void *stream);

/* ---- timing helper -----------------------------------------------------------------
 * Runs b2s_ntt `iters` times back to back on `stream` bracketed by CUDA events and
 * returns the mean milliseconds per call in *ms (used by bench.py for the roofline of the
 * dominant kernel; torch.cuda.Event only sees torch's current stream). */
int b2s_ntt_timed(const uint64_t *d_in, uint64_t in_stride, uint32_t n_in, uint64_t *d_out, uint64_t out_stride,
                  uint32_t log_n, uint32_t n_planes, uint64_t omega, uint64_t offset, int inverse, void *stream,
                  uint32_t iters, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H */
