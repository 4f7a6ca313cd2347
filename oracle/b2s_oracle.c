/*
 * oracle/b2s_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the algorithms on the hot path of
 * aszepieniec/stark-brainfuck (pure Python), used ONLY as the checker in tests/,
 * in __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference arm.
 * Nothing under stark_brainfuck_b200/ imports, links or executes this file.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function here against golden
 * vectors produced by running the unmodified reference (tests/golden/make_golden.py)
 * and against CPython's own hashlib.blake2b / pickle.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/).  Third-party algorithms the reference calls but does not contain:
 *   - BLAKE2b-512 (CPython 3.12 hashlib.blake2b, unkeyed, 64-byte digest): restated
 *     from RFC 7693.
 *   - pickle protocol 4 integer/frame encoding (CPython 3.12 Modules/_pickle.c
 *     save_long / frame header): restated from PEP 3154 + pickletools docs.
 *
 * Layout: field elements are canonical u64 in [0,p); extension-field vectors are three
 * u64 planes (c0,c1,c2) of n elements each, plane stride given explicitly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL /* code/algebra.py:114 */

/* ---------------------------------------------------------------- base field */
/* code/algebra.py:92-93 */
u64 orc_add(u64 a, u64 b) { return (u64)(((u128)a + b) % GL_P); }
/* code/algebra.py:95-96 */
u64 orc_sub(u64 a, u64 b) { return (u64)(((u128)GL_P + a - b) % GL_P); }
/* code/algebra.py:98-99 */
u64 orc_neg(u64 a) { return (GL_P - a) % GL_P; }
/* code/algebra.py:89-90 */
u64 orc_mul(u64 a, u64 b) { return (u64)(((u128)a * b) % GL_P); }

/* code/algebra.py:39-46: MSB-first square and multiply */
u64 orc_pow(u64 base, u64 e) {
    u64 acc = 1;
    for (int i = 63; i >= 0; --i) {
        acc = orc_mul(acc, acc);
        if ((e >> i) & 1) acc = orc_mul(acc, base);
    }
    return acc;
}

/* code/algebra.py:101-103 computes the inverse with xgcd; the inverse is unique, so
 * Fermat (a^(p-2)) returns the same canonical value, including inverse(0) == 0. */
u64 orc_inv(u64 a) { return orc_pow(a, GL_P - 2); }

/* code/algebra.py:138-142: big-endian bytes mod p */
u64 orc_sample(const u8 *bytes, u32 len) {
    u128 acc = 0;
    for (u32 i = 0; i < len; ++i) acc = ((acc << 8) | bytes[i]) % GL_P;
    return (u64)acc;
}

/* ---------------------------------------------------------------- NTT */
/* code/ntt.py:4-23: recursive radix-2 decimation in time, natural order in and out.
 * in[] is read with a stride (values[1::2] / values[::2] of the reference), tw[] holds
 * the powers of the TOP-level root, tw_stride selects the powers of omega^(2^level). */
static void ntt_rec(const u64 *in, size_t stride, size_t n, u64 *out, const u64 *tw, size_t tw_stride) {
    if (n == 1) { /* code/ntt.py:8-9 */
        out[0] = in[0];
        return;
    }
    size_t half = n / 2;
    ntt_rec(in + stride, 2 * stride, half, out + half, tw, 2 * tw_stride); /* odds, code/ntt.py:20 */
    ntt_rec(in, 2 * stride, half, out, tw, 2 * tw_stride);                 /* evens, code/ntt.py:21 */
    for (size_t i = 0; i < half; ++i) {                                    /* code/ntt.py:23 */
        u64 e = out[i];
        u64 t = orc_mul(tw[i * tw_stride], out[half + i]);
        out[i] = orc_add(e, t);        /* i < half:  evens[i] + w^i * odds[i]           */
        out[half + i] = orc_sub(e, t); /* i >= half: w^(i+half) = -w^i for a primitive root */
    }
}

/* returns 0 on success; -1 if n is not a power of two; -2 if omega^n != 1;
 * -3 if omega^(n/2) == 1 (the three asserts of code/ntt.py:5-6, :13-16) */
int orc_ntt(u64 omega, const u64 *in, u64 *out, u64 n) {
    if (n & (n - 1)) return -1;
    if (n <= 1) {
        if (n == 1) out[0] = in[0];
        return 0;
    }
    if (orc_pow(omega, n) != 1) return -2;
    if (orc_pow(omega, n / 2) == 1) return -3;
    u64 *tw = (u64 *)malloc(sizeof(u64) * (n / 2));
    if (!tw) return -9;
    tw[0] = 1;
    for (u64 i = 1; i < n / 2; ++i) tw[i] = orc_mul(tw[i - 1], omega);
    ntt_rec(in, 1, n, out, tw, 1);
    free(tw);
    return 0;
}

/* code/ntt.py:26-42: n^-1 * ntt(omega^-1, values) */
int orc_intt(u64 omega, const u64 *in, u64 *out, u64 n) {
    if (n & (n - 1)) return -1;
    if (orc_pow(omega, n) != 1) return -2;
    if (n == 1) {
        out[0] = in[0];
        return 0;
    }
    if (n == 0) return 0;
    if (orc_pow(omega, n / 2) == 1) return -3;
    u64 ninv = orc_inv(n % GL_P); /* code/ntt.py:39 */
    int rc = orc_ntt(orc_inv(omega), in, out, n);
    if (rc) return rc;
    for (u64 i = 0; i < n; ++i) out[i] = orc_mul(ninv, out[i]);
    return 0;
}

/* code/univariate.py:168-169: c_i <- factor^i * c_i (base-field factor applied to one plane) */
void orc_scale(u64 factor, const u64 *in, u64 *out, u64 n) {
    u64 f = 1;
    for (u64 i = 0; i < n; ++i) {
        out[i] = orc_mul(f, in[i]);
        f = orc_mul(f, factor);
    }
}

/* code/ntt.py:164-168 and code/fri.py:26-30: scale by offset, zero-pad to n, ntt */
int orc_coset_evaluate(u64 offset, u64 omega, const u64 *coeffs, u64 m, u64 *out, u64 n) {
    if (m > n) return -4;
    u64 *tmp = (u64 *)calloc(n ? n : 1, sizeof(u64));
    if (!tmp) return -9;
    orc_scale(offset, coeffs, tmp, m);
    int rc = orc_ntt(omega, tmp, out, n);
    free(tmp);
    return rc;
}

/* code/ntt.py:171-174: intt, then scale by offset^-1; keeps all n coefficients */
int orc_coset_interpolate(u64 offset, u64 omega, const u64 *values, u64 *out, u64 n) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * (n ? n : 1));
    if (!tmp) return -9;
    int rc = orc_intt(omega, values, tmp, n);
    if (!rc) orc_scale(orc_inv(offset), tmp, out, n);
    free(tmp);
    return rc;
}

/* ---------------------------------------------------------------- cubic extension */
/* F_p[X]/(X^3 - X + 1): code/extension_field.py:88-98.  Elements are (c0,c1,c2). */

/* code/extension_field.py:65-66: schoolbook product, then reduce with X^3 = X - 1, X^4 = X^2 - X */
void orc_xmul(const u64 a[3], const u64 b[3], u64 r[3]) {
    u64 c0 = orc_mul(a[0], b[0]);
    u64 c1 = orc_add(orc_mul(a[0], b[1]), orc_mul(a[1], b[0]));
    u64 c2 = orc_add(orc_add(orc_mul(a[0], b[2]), orc_mul(a[1], b[1])), orc_mul(a[2], b[0]));
    u64 c3 = orc_add(orc_mul(a[1], b[2]), orc_mul(a[2], b[1]));
    u64 c4 = orc_mul(a[2], b[2]);
    r[0] = orc_sub(c0, c3);
    r[1] = orc_sub(orc_add(c1, c3), c4);
    r[2] = orc_add(c2, c4);
}
void orc_xadd(const u64 a[3], const u64 b[3], u64 r[3]) { /* code/extension_field.py:68-69 */
    for (int i = 0; i < 3; ++i) r[i] = orc_add(a[i], b[i]);
}
void orc_xsub(const u64 a[3], const u64 b[3], u64 r[3]) { /* code/extension_field.py:71-72 */
    for (int i = 0; i < 3; ++i) r[i] = orc_sub(a[i], b[i]);
}

/* code/extension_field.py:77-81 inverts with a polynomial xgcd against the modulus; the
 * inverse is unique, so solving M(a)·x = (1,0,0) with Cramer's rule, where M(a) is the
 * multiplication-by-a matrix in the basis 1,X,X^2, returns the same triple.  inverse(0)
 * is not defined by the reference (the xgcd divides by a zero leading coefficient). */
void orc_xinv(const u64 a[3], u64 r[3]) {
    /* columns of M: a*1, a*X, a*X^2 */
    u64 e1[3] = {0, 1, 0}, e2[3] = {0, 0, 1};
    u64 c0[3] = {a[0], a[1], a[2]}, c1[3], c2[3];
    orc_xmul(a, e1, c1);
    orc_xmul(a, e2, c2);
    /* M = [c0 c1 c2] (columns).  x = adj(M)[:,0] / det: first column of the inverse */
    u64 m00 = c0[0], m01 = c1[0], m02 = c2[0];
    u64 m10 = c0[1], m11 = c1[1], m12 = c2[1];
    u64 m20 = c0[2], m21 = c1[2], m22 = c2[2];
    u64 A00 = orc_sub(orc_mul(m11, m22), orc_mul(m12, m21));
    u64 A10 = orc_sub(orc_mul(m12, m20), orc_mul(m10, m22));
    u64 A20 = orc_sub(orc_mul(m10, m21), orc_mul(m11, m20));
    u64 det = orc_add(orc_add(orc_mul(m00, A00), orc_mul(m01, A10)), orc_mul(m02, A20));
    u64 dinv = orc_inv(det);
    r[0] = orc_mul(A00, dinv);
    r[1] = orc_mul(A10, dinv);
    r[2] = orc_mul(A20, dinv);
}
/* code/extension_field.py:83-86 */
void orc_xdiv(const u64 a[3], const u64 b[3], u64 r[3]) {
    u64 bi[3];
    orc_xinv(b, bi);
    orc_xmul(a, bi, r);
}

/* code/extension_field.py:100-111: three big-endian chunks of len/3 bytes each */
void orc_xsample(const u8 *bytes, u32 len, u64 r[3]) {
    u32 chunk = len / 3;
    for (int i = 0; i < 3; ++i) r[i] = orc_sample(bytes + i * chunk, chunk);
}

/* XFE vectors are three planes; a lifted base-field root acts on each plane separately
 * (code/fri.py:37 lifts omega; all 2-power roots of unity of F_p^3 lie in F_p). */
int orc_xntt(u64 omega, const u64 *in, u64 in_stride, u64 *out, u64 out_stride, u64 n, int inverse) {
    for (int pl = 0; pl < 3; ++pl) {
        int rc = inverse ? orc_intt(omega, in + pl * in_stride, out + pl * out_stride, n)
                         : orc_ntt(omega, in + pl * in_stride, out + pl * out_stride, n);
        if (rc) return rc;
    }
    return 0;
}

/* code/univariate.py:168-169 with an extension-field factor: c_i <- factor^i * c_i */
void orc_xscale(const u64 factor[3], const u64 *in, u64 in_stride, u64 *out, u64 out_stride, u64 n) {
    u64 f[3] = {1, 0, 0};
    for (u64 i = 0; i < n; ++i) {
        u64 c[3] = {in[i], in[in_stride + i], in[2 * in_stride + i]}, r[3], nf[3];
        orc_xmul(f, c, r);
        out[i] = r[0];
        out[out_stride + i] = r[1];
        out[2 * out_stride + i] = r[2];
        orc_xmul(f, factor, nf);
        f[0] = nf[0]; f[1] = nf[1]; f[2] = nf[2];
    }
}

/* code/univariate.py:145-154: evaluate with a running power of the point (not Horner),
 * base-field polynomial at base-field points. */
void orc_eval_points(const u64 *coeffs, u64 m, const u64 *points, u64 *out, u64 npoints) {
    for (u64 k = 0; k < npoints; ++k) {
        u64 xi = 1, val = 0;
        for (u64 i = 0; i < m; ++i) {
            val = orc_add(val, orc_mul(coeffs[i], xi));
            xi = orc_mul(xi, points[k]);
        }
        out[k] = val;
    }
}
/* same with extension-field coefficients and extension-field points (planes) */
void orc_xeval_points(const u64 *coeffs, u64 cstride, u64 m, const u64 *points, u64 pstride, u64 *out,
                      u64 ostride, u64 npoints) {
    for (u64 k = 0; k < npoints; ++k) {
        u64 x[3] = {points[k], points[pstride + k], points[2 * pstride + k]};
        u64 xi[3] = {1, 0, 0}, val[3] = {0, 0, 0};
        for (u64 i = 0; i < m; ++i) {
            u64 c[3] = {coeffs[i], coeffs[cstride + i], coeffs[2 * cstride + i]}, t[3], nx[3];
            orc_xmul(c, xi, t);
            orc_xadd(val, t, val);
            orc_xmul(xi, x, nx);
            xi[0] = nx[0]; xi[1] = nx[1]; xi[2] = nx[2];
        }
        out[k] = val[0];
        out[ostride + k] = val[1];
        out[2 * ostride + k] = val[2];
    }
}

/* ---------------------------------------------------------------- FRI fold */
/* code/fri.py:127-128:
 *   c'[i] = 2^-1 * ((1 + alpha/(offset*omega^i)) * c[i] + (1 - alpha/(offset*omega^i)) * c[N/2+i])
 * written with the same operations in the same order (the divisor is a lifted base element). */
void orc_fri_fold(const u64 *cw, u64 stride, u64 N, const u64 alpha[3], u64 offset, u64 omega, u64 *out,
                  u64 ostride) {
    u64 one[3] = {1, 0, 0}, two[3] = {2, 0, 0}, half[3];
    orc_xinv(two, half);
    u64 wi = 1;
    for (u64 i = 0; i < N / 2; ++i) {
        u64 x[3] = {orc_mul(offset, wi), 0, 0}, q[3], l[3], r[3], s[3], t[3];
        orc_xdiv(alpha, x, q);
        orc_xadd(one, q, l);
        orc_xsub(one, q, r);
        u64 a[3] = {cw[i], cw[stride + i], cw[2 * stride + i]};
        u64 b[3] = {cw[N / 2 + i], cw[stride + N / 2 + i], cw[2 * stride + N / 2 + i]};
        orc_xmul(l, a, s);
        orc_xmul(r, b, t);
        orc_xadd(s, t, s);
        orc_xmul(half, s, t);
        out[i] = t[0];
        out[ostride + i] = t[1];
        out[2 * ostride + i] = t[2];
        wi = orc_mul(wi, omega);
    }
}

/* ---------------------------------------------------------------- quotient codewords */
/* code/multivariate.py:105-116 MPolynomial.evaluate: acc = sum over monomials of
 * coefficient * prod_i point[i]^k[i], in the order of the flattened dictionary. */
static void xpow_small(const u64 a[3], u32 e, u64 r[3]) {
    u64 acc[3] = {1, 0, 0}, t[3];
    for (u32 k = 0; k < e; ++k) {
        orc_xmul(acc, a, t);
        memcpy(acc, t, sizeof(t));
    }
    memcpy(r, acc, sizeof(acc));
}

/* code/table.py:155-178, :190-236, :253-286 and code/permutation_argument.py:11-20 for one table:
 * out[c][i] = mpo_c.evaluate(point_i) * lift(zerofier_inverse[i]) over the domain offset*omega^i.
 * cw: `width` extension-field codewords of N values (planes c0,c1,c2 of codeword v at 3v, 3v+1, 3v+2);
 * variables >= width read row (i + shift) mod N.  zerofier kinds: 1 boundary (x - 1)^-1,
 * 2 transition (x^height - 1)^-1 (x - omicron_inv) or 0 when height == 0, 3 terminal (x - omicron_inv)^-1.
 * Returns 1 if a zerofier vanishes on the domain (the reference's batch_inverse asserts), else 0. */
int orc_quotients(const u64 *cw, u64 N, u32 width, u64 shift, u32 n_constraints, const u32 *mono_off,
                  const u64 *coeffs, const u32 *factors, u32 max_factors, u32 kind, u64 height, u64 omicron_inv,
                  u64 offset, u64 omega, u64 *out) {
    int flag = 0;
    u64 wi = 1;
    for (u64 i = 0; i < N; ++i) {
        const u64 x = orc_mul(offset, wi);
        u64 zinv;
        if (kind == 1) {
            const u64 z = orc_sub(x, 1);
            flag |= z == 0;
            zinv = orc_inv(z);
        } else if (kind == 2) {
            if (height == 0) {
                zinv = 0;
            } else {
                const u64 z = orc_sub(orc_pow(x, height), 1);
                flag |= z == 0;
                zinv = orc_mul(orc_inv(z), orc_sub(x, omicron_inv));
            }
        } else {
            const u64 z = orc_sub(x, omicron_inv);
            flag |= z == 0;
            zinv = orc_inv(z);
        }
        const u64 inext = (i + shift) % N;
        for (u32 c = 0; c < n_constraints; ++c) {
            u64 acc[3] = {0, 0, 0};
            for (u32 m = mono_off[c]; m < mono_off[c + 1]; ++m) {
                u64 prod[3] = {coeffs[3 * m], coeffs[3 * m + 1], coeffs[3 * m + 2]}, t[3], pw[3];
                for (u32 f = 0; f < max_factors; ++f) {
                    const u32 fac = factors[m * max_factors + f], e = fac & 0xFF;
                    if (!e) continue;
                    u32 v = fac >> 8;
                    u64 at = i;
                    if (v >= width) {
                        v -= width;
                        at = inext;
                    }
                    const u64 xv[3] = {cw[(3 * (u64)v) * N + at], cw[(3 * (u64)v + 1) * N + at],
                                       cw[(3 * (u64)v + 2) * N + at]};
                    xpow_small(xv, e, pw);
                    orc_xmul(prod, pw, t);
                    memcpy(prod, t, sizeof(t));
                }
                orc_xadd(acc, prod, t);
                memcpy(acc, t, sizeof(t));
            }
            const u64 zl[3] = {zinv, 0, 0};
            u64 q[3];
            orc_xmul(acc, zl, q);
            out[(3 * (u64)c) * N + i] = q[0];
            out[(3 * (u64)c + 1) * N + i] = q[1];
            out[(3 * (u64)c + 2) * N + i] = q[2];
        }
        wi = orc_mul(wi, omega);
    }
    return flag;
}

/* code/brainfuck_stark.py:241-298, the nonlinear combination: terms c and x^shift * c of every
 * codeword (:247-291; lift of base-field values, code/extension_field.py:113-116), weighted sum
 * (:298).  out[j] = sum_c wa_c * col_c[j] + wb_c * (x_j^shift_c * col_c[j]), x_j = offset*omega^j.
 * cols[c]: `planes[c]` (1 or 3) planes of N values, strides[c] apart; a zero wb means the column
 * has no shifted term (the randomizer codeword, :242). */
void orc_combination(const u64 *const *cols, const u64 *strides, const u32 *planes, const u64 *wa, const u64 *wb,
                     const u64 *shifts, u32 n_cols, u64 N, u64 offset, u64 omega, u64 *out, u64 out_stride) {
    u64 wi = 1;
    for (u64 j = 0; j < N; ++j) {
        const u64 x = orc_mul(offset, wi);
        u64 acc[3] = {0, 0, 0}, t[3], u[3];
        for (u32 c = 0; c < n_cols; ++c) {
            u64 v[3] = {cols[c][j], 0, 0};
            if (planes[c] == 3) {
                v[1] = cols[c][strides[c] + j];
                v[2] = cols[c][2 * strides[c] + j];
            }
            orc_xmul(wa + 3 * c, v, t); /* weight * unshifted term */
            orc_xadd(acc, t, u);
            memcpy(acc, u, sizeof(u));
            if (wb[3 * c] | wb[3 * c + 1] | wb[3 * c + 2]) {
                const u64 xs[3] = {orc_pow(x, shifts[c]), 0, 0};
                u64 sv[3];
                orc_xmul(xs, v, sv); /* lift(x^shift) * c[j] */
                orc_xmul(wb + 3 * c, sv, t);
                orc_xadd(acc, t, u);
                memcpy(acc, u, sizeof(u));
            }
        }
        out[j] = acc[0];
        out[out_stride + j] = acc[1];
        out[2 * out_stride + j] = acc[2];
        wi = orc_mul(wi, omega);
    }
}

/* ---------------------------------------------------------------- BLAKE2b-512 (RFC 7693) */
static const u64 B2B_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                              0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                              0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const u8 B2B_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

static inline u64 rotr64(u64 x, int r) { return (x >> r) | (x << (64 - r)); }

static void b2b_compress(u64 h[8], const u8 block[128], u128 t, int last) {
    u64 v[16], m[16];
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = B2B_IV[i]; }
    v[12] ^= (u64)t;
    v[13] ^= (u64)(t >> 64);
    if (last) v[14] = ~v[14];
    for (int i = 0; i < 16; ++i) {
        u64 w = 0;
        for (int j = 7; j >= 0; --j) w = (w << 8) | block[8 * i + j];
        m[i] = w;
    }
#define B2B_G(a, b, c, d, x, y)                                                          \
    do {                                                                                 \
        v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32);                        \
        v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 24);                        \
        v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16);                        \
        v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 63);                        \
    } while (0)
    for (int r = 0; r < 12; ++r) {
        const u8 *s = B2B_SIGMA[r];
        B2B_G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        B2B_G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        B2B_G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        B2B_G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        B2B_G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        B2B_G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        B2B_G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        B2B_G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
#undef B2B_G
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

/* hashlib.blake2b(msg).digest(): unkeyed, digest_size 64 (code/merkle.py:31-32, :38-39) */
void orc_blake2b(const u8 *msg, u64 len, u8 out[64]) {
    u64 h[8];
    for (int i = 0; i < 8; ++i) h[i] = B2B_IV[i];
    h[0] ^= 0x01010040ULL; /* digest length 64, no key, fanout 1, depth 1 */
    u64 off = 0;
    while (len - off > 128) {
        b2b_compress(h, msg + off, (u128)(off + 128), 0);
        off += 128;
    }
    u8 last[128];
    memset(last, 0, 128);
    memcpy(last, msg + off, len - off);
    b2b_compress(h, last, (u128)len, 1);
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) out[8 * i + j] = (u8)(h[i] >> (8 * j));
}

/* ---------------------------------------------------------------- pickle leaf preimages */
/* Same struct layout as include/b2s.h (declared independently on purpose). */
#define ORC_TPL_MAX_BYTES 2048
typedef struct {
    u32 n_slots;        /* 1: BaseFieldElement leaves, 3: ExtensionFieldElement leaves */
    u32 trim;           /* 1: template index = #coefficients after trimming trailing zeros (code/extension_field.py:6-9) */
    u32 seg_off[4][5];  /* template k: segments j=0..k are bytes[seg_off[k][j] .. seg_off[k][j+1]) */
    u8 bytes[ORC_TPL_MAX_BYTES];
} orc_leaf_templates;

/* CPython 3.12 _pickle.c save_long for a non-negative int < 2^64, protocol 4:
 * BININT1 'K', BININT2 'M', BININT 'J' (< 2^31), else LONG1 0x8a with
 * nbytes = (bit_length >> 3) + 1 little-endian two's complement bytes. */
u32 orc_pickle_uint(u64 v, u8 *out) {
    if (v < 256) { out[0] = 0x4b; out[1] = (u8)v; return 2; }
    if (v < 65536) { out[0] = 0x4d; out[1] = (u8)v; out[2] = (u8)(v >> 8); return 3; }
    if (v < 0x80000000ULL) {
        out[0] = 0x4a;
        for (int i = 0; i < 4; ++i) out[1 + i] = (u8)(v >> (8 * i));
        return 5;
    }
    u32 bits = 64 - (u32)__builtin_clzll(v);
    u32 nb = (bits >> 3) + 1;
    out[0] = 0x8a;
    out[1] = (u8)nb;
    for (u32 i = 0; i < nb; ++i) out[2 + i] = i < 8 ? (u8)(v >> (8 * i)) : 0;
    return 2 + nb;
}

/* pickle.dumps(leaf) of code/merkle.py:30 for a field-element leaf: PROTO 4, FRAME(len), body */
u32 orc_leaf_preimage(const orc_leaf_templates *tp, const u64 *c, u8 *out) {
    u32 k = tp->n_slots;
    if (tp->trim)
        while (k > 0 && c[k - 1] == 0) --k;
    u8 *p = out + 11;
    for (u32 j = 0; j <= k; ++j) {
        u32 a = tp->seg_off[k][j], b = tp->seg_off[k][j + 1];
        memcpy(p, tp->bytes + a, b - a);
        p += b - a;
        if (j < k) p += orc_pickle_uint(c[j], p);
    }
    u64 body = (u64)(p - (out + 11));
    out[0] = 0x80; out[1] = 0x04; out[2] = 0x95;
    for (int i = 0; i < 8; ++i) out[3 + i] = (u8)(body >> (8 * i));
    return (u32)(11 + body);
}

/* ---------------------------------------------------------------- Merkle */
/* code/merkle.py:8-41.  nodes: 2*npo2 slots of 64 bytes, heap layout, nodes[1] = root,
 * nodes[npo2+i] = leaf digest i.  Slot 0 is left zeroed (the reference stores a value
 * there that nothing reads).  Field-element leaves always come in power-of-two counts. */
static void merkle_upper(u8 *nodes, u64 npo2) {
    for (u64 k = npo2 - 1; k >= 1; --k) /* code/merkle.py:35-41 */
        orc_blake2b(nodes + 128 * k, 128, nodes + 64 * k);
}

/* code/merkle.py:35-41 for callers that already hold the leaf-level digests in slots [npo2, 2 npo2) */
void orc_merkle_upper(u8 *nodes, u64 npo2) { merkle_upper(nodes, npo2); }

int orc_merkle_field(const orc_leaf_templates *tp, const u64 *planes, u64 stride, u64 n, u8 *nodes) {
    if (n == 0 || (n & (n - 1))) return -1;
    memset(nodes, 0, 64);
    u8 buf[640];
    for (u64 i = 0; i < n; ++i) { /* code/merkle.py:29-32 */
        u64 c[3] = {0, 0, 0};
        for (u32 s = 0; s < tp->n_slots; ++s) c[s] = planes[s * stride + i];
        u32 len = orc_leaf_preimage(tp, c, buf);
        orc_blake2b(buf, len, nodes + 64 * (n + i));
    }
    merkle_upper(nodes, n);
    return 0;
}

/* Arbitrary picklable leaves (code/test_merkle.py:57-61): the caller pickles, we hash.
 * Non-power-of-two counts keep the reference's 32-byte zero placeholders
 * (code/merkle.py:26), so first-level parents may hash 96- or 64-byte messages. */
int orc_merkle_blobs(const u8 *bytes, const u64 *offsets, u64 n, u64 npo2, u8 *nodes) {
    if (n == 0 || npo2 < n || (npo2 & (npo2 - 1))) return -1;
    memset(nodes, 0, 64 * 2 * npo2);
    for (u64 i = 0; i < n; ++i)
        orc_blake2b(bytes + offsets[i], offsets[i + 1] - offsets[i], nodes + 64 * (npo2 + i));
    if (npo2 == 1) return 0;
    for (u64 k = npo2 - 1; k >= npo2 / 2; --k) {
        u8 msg[128];
        u32 len = 0;
        for (u64 c = 2 * k; c <= 2 * k + 1; ++c) {
            u32 l = (c - npo2 < n) ? 64 : 32;
            if (l == 64) memcpy(msg + len, nodes + 64 * c, 64); else memset(msg + len, 0, 32);
            len += l;
        }
        orc_blake2b(msg, len, nodes + 64 * k);
        if (k == 1) return 0;
    }
    for (u64 k = npo2 / 2 - 1; k >= 1; --k) orc_blake2b(nodes + 128 * k, 128, nodes + 64 * k);
    return 0;
}

/* code/salted_merkle.py:25-35 over rows of codewords (code/brainfuck_stark.py:178-180, :197-199):
 * leaf r = blake2b(pickle.dumps(row_r) | pickle.dumps(salt_r)).  The row pickle is the byte template of the
 * caller's tuple with one integer per emitted plane (mode 0/1; mode 2 = a trimmed coefficient, must be zero,
 * code/extension_field.py:6-9).  Rows whose shape differs from the template are listed in `exceptions` and
 * left unhashed.  rows == NULL: all n rows.  Returns the number of exceptions, or -1 on bad arguments. */
long orc_row_leaves(const u64 *const *planes, const u8 *modes, u32 n_planes, u64 n, const u8 *tpl, const u32 *seg_off,
                    u32 n_slots, const u8 *salts, u32 salt_len, const u8 *salt_pre, u32 salt_pre_len,
                    const u8 *salt_suf, u32 salt_suf_len, const u32 *rows, u64 n_rows, u8 *nodes, u32 *exceptions) {
    if (n == 0 || (n & (n - 1))) return -1;
    const u64 count = rows ? n_rows : n;
    const u32 tpl_len = seg_off[n_slots + 1];
    u8 *msg = (u8 *)malloc(11 + tpl_len + 11 * (size_t)n_planes + salt_pre_len + salt_len + salt_suf_len + 16);
    long n_exc = 0;
    for (u64 t = 0; t < count; ++t) {
        const u64 r = rows ? rows[t] : t;
        int ok = 1;
        for (u32 p = 0; p < n_planes; ++p) {
            const u64 v = planes[p][r];
            if (modes[p] == 2 ? v != 0 : (modes[p] == 1 && v == 0)) ok = 0;
        }
        if (!ok) {
            exceptions[n_exc++] = (u32)r;
            continue;
        }
        u32 len = 11, slot = 0;
        for (u32 p = 0; p < n_planes; ++p) {
            if (modes[p] == 2) continue;
            memcpy(msg + len, tpl + seg_off[slot], seg_off[slot + 1] - seg_off[slot]);
            len += seg_off[slot + 1] - seg_off[slot];
            len += orc_pickle_uint(planes[p][r], msg + len);
            ++slot;
        }
        if (slot != n_slots) {
            free(msg);
            return -1;
        }
        memcpy(msg + len, tpl + seg_off[slot], seg_off[slot + 1] - seg_off[slot]);
        len += seg_off[slot + 1] - seg_off[slot];
        const u64 body = len - 11;
        msg[0] = 0x80;
        msg[1] = 0x04;
        msg[2] = 0x95;
        for (int i = 0; i < 8; ++i) msg[3 + i] = (u8)(body >> (8 * i));
        if (salts) {
            memcpy(msg + len, salt_pre, salt_pre_len);
            len += salt_pre_len;
            memcpy(msg + len, salts + r * salt_len, salt_len);
            len += salt_len;
            memcpy(msg + len, salt_suf, salt_suf_len);
            len += salt_suf_len;
        }
        orc_blake2b(msg, len, nodes + 64 * (n + r));
    }
    free(msg);
    return n_exc;
}

/* code/merkle.py:46-52: sibling digests from the leaf level upward */
void orc_merkle_open(const u8 *nodes, u64 npo2, u64 index, u8 *path /* depth*64 */) {
    u64 k = npo2 | index;
    u32 j = 0;
    while (k > 1) {
        memcpy(path + 64 * j, nodes + 64 * (k ^ 1), 64);
        ++j;
        k >>= 1;
    }
}
