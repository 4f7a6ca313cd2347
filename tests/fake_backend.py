"""TEST-ONLY stand-in for libb2s.so: the same C-ABI entry points (include/b2s.h) executed by the
CPU oracle on HOST memory, so that the host glue (marshalling, identity rules, FRI round loop,
drop-in patching) can be exercised in a container without a GPU.

It is injected explicitly (`Engine(lib=FakeLib(), device="cpu")`) by tests; nothing in
stark_brainfuck_b200/ knows about it and the product path never falls back to it.
"""
import ctypes as C

import numpy as np

from oracle import oracle as orc

ERR_ROOT, ERR_PRIM = -11, -12


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if hasattr(p, "_obj"):  # byref(x)
        return C.addressof(p._obj)
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return C.cast(p, C.c_void_p).value or 0


def _set_i32(p, v):
    C.c_int32.from_address(_addr(p)).value = int(v)


def _u64(addr, count):
    return np.ctypeslib.as_array((C.c_uint64 * count).from_address(addr))


def _u8(addr, count):
    return np.ctypeslib.as_array((C.c_uint8 * count).from_address(addr))


def _otpl(tp):
    src = tp._obj if hasattr(tp, "_obj") else tp.contents
    dst = orc.LeafTemplates()
    assert C.sizeof(dst) == C.sizeof(src)
    C.memmove(C.byref(dst), C.byref(src), C.sizeof(src))
    return dst


class FakeLib:
    def __init__(self):
        self.launches = 0
        self.err = b""

    def b2s_last_error(self):
        return self.err

    def b2s_launch_count(self):
        return self.launches

    def b2s_ntt(self, d_in, in_stride, n_in, d_out, out_stride, log_n, q, omega, offset, inverse, stream):
        n = 1 << log_n
        P = orc.P
        if pow(omega, n, P) != 1:
            self.err = b"primitive root must be nth root of unity"
            return ERR_ROOT
        if log_n >= 1 and pow(omega, n // 2, P) == 1:
            self.err = b"primitive root is not primitive nth root of unity"
            return ERR_PRIM
        self.launches += 1
        ins = [_u64(_addr(d_in) + 8 * in_stride * i, n_in).copy() for i in range(q)]
        for i in range(q):
            out = _u64(_addr(d_out) + 8 * out_stride * i, n)
            if inverse:
                out[:] = orc.coset_interpolate(offset, omega, ins[i]) if n > 1 else ins[i]
            else:
                out[:] = orc.coset_evaluate(offset, omega, ins[i], n) if n > 1 else ins[i]
        return 0

    def b2s_ntt_host(self, h_in, in_stride, n_in, h_out, out_stride, log_n, q, omega, offset, inverse):
        return self.b2s_ntt(h_in, in_stride, n_in, h_out, out_stride, log_n, q, omega, offset, inverse, None)

    def b2s_scale(self, d_in, in_stride, d_out, out_stride, n, q, factor, stream):
        self.launches += 1
        if q == 1:
            _u64(_addr(d_out), n)[:] = orc.scale(int(factor[0]), _u64(_addr(d_in), n))
        else:
            x = np.stack([_u64(_addr(d_in) + 8 * in_stride * i, n) for i in range(3)])
            r = orc.xscale([int(factor[i]) for i in range(3)], x)
            for i in range(3):
                _u64(_addr(d_out) + 8 * out_stride * i, n)[:] = r[i]
        return 0

    def b2s_eval_points(self, d_c, cstride, cq, m, d_p, pstride, pq, k, d_out, ostride, stream):
        self.launches += 1
        c = np.zeros((3, m), dtype=np.uint64)
        p = np.zeros((3, k), dtype=np.uint64)
        for i in range(cq):
            c[i] = _u64(_addr(d_c) + 8 * cstride * i, m)
        for i in range(pq):
            p[i] = _u64(_addr(d_p) + 8 * pstride * i, k)
        if cq == 1 and pq == 1:
            _u64(_addr(d_out), k)[:] = orc.eval_points(c[0], p[0])
        else:
            r = orc.eval_points(c, p)
            for i in range(3):
                _u64(_addr(d_out) + 8 * ostride * i, k)[:] = r[i]
        return 0

    def b2s_merkle_field(self, d_planes, stride, n, tpl, d_nodes, stream):
        self.launches += 1
        t = _otpl(tpl)
        planes = np.stack([_u64(_addr(d_planes) + 8 * stride * i, n) for i in range(t.n_slots)])
        nodes = orc.merkle_field(t, planes if t.n_slots == 3 else planes[0])
        _u8(_addr(d_nodes), 128 * n)[:] = nodes.reshape(-1)
        return 0

    def b2s_merkle_blobs(self, d_bytes, d_offsets, n, npo2, d_nodes, stream):
        self.launches += 1
        offs = _u64(_addr(d_offsets), n + 1)
        data = _u8(_addr(d_bytes), int(offs[n])) if offs[n] else np.zeros(0, dtype=np.uint8)
        blobs = [bytes(data[int(offs[i]):int(offs[i + 1])]) for i in range(n)]
        _u8(_addr(d_nodes), 128 * npo2)[:] = orc.merkle_blobs(blobs).reshape(-1)
        return 0

    def b2s_merkle_rows(self, h_planes, h_modes, n_planes, n, h_tpl, h_seg_off, n_slots, d_salts, salt_len, h_pre,
                        pre_len, h_suf, suf_len, d_rows, n_rows, d_nodes, build_upper, h_exc, h_n_exc, stream):
        self.launches += 1
        ptrs = _u64(_addr(h_planes), n_planes)
        planes = [_u64(int(ptrs[p]), n) for p in range(n_planes)]
        modes = _u8(_addr(h_modes), n_planes)
        seg = np.ctypeslib.as_array((C.c_uint32 * (n_slots + 2)).from_address(_addr(h_seg_off)))
        tpl = bytes(_u8(_addr(h_tpl), int(seg[n_slots + 1]))) if seg[n_slots + 1] else b""
        salts = _u8(_addr(d_salts), n * salt_len).reshape(n, salt_len) if _addr(d_salts) else None
        rows = (np.ctypeslib.as_array((C.c_uint32 * n_rows).from_address(_addr(d_rows))) if _addr(d_rows) else None)
        nodes = _u8(_addr(d_nodes), 128 * n).reshape(2 * n, 64)
        if rows is None:
            nodes[0] = 0
        exc = orc.row_leaves(planes, modes, tpl, seg, n, nodes, salts, bytes(_u8(_addr(h_pre), pre_len)) if pre_len else b"",
                             bytes(_u8(_addr(h_suf), suf_len)) if suf_len else b"", rows)
        if len(exc):
            np.ctypeslib.as_array((C.c_uint32 * len(exc)).from_address(_addr(h_exc)))[:] = exc
        _set_i32(h_n_exc, len(exc))
        if build_upper and not len(exc) and n > 1:
            nodes[:] = orc.merkle_upper(nodes.copy())
        return 0

    def b2s_merkle_upper(self, d_nodes, npo2, stream):
        self.launches += 1
        view = _u8(_addr(d_nodes), 128 * npo2).reshape(-1, 64)
        view[:] = orc.merkle_upper(view.copy())
        return 0

    def b2s_merkle_open(self, d_nodes, npo2, h_idx, n_idx, h_paths, stream):
        self.launches += 1
        nodes = _u8(_addr(d_nodes), 128 * npo2).reshape(-1, 64)
        idx = _u64(_addr(h_idx), n_idx)
        depth = npo2.bit_length() - 1
        out = _u8(_addr(h_paths), n_idx * depth * 64).reshape(n_idx, depth, 64)
        for q in range(n_idx):
            for j, b in enumerate(orc.merkle_open(nodes, int(idx[q]))):
                out[q, j] = np.frombuffer(b, dtype=np.uint8)
        return 0

    def b2s_fri_fold(self, d_cw, cw_stride, N, alpha, offset, omega, d_next, next_stride, tpl, d_next_nodes, stream):
        self.launches += 1
        cw = np.stack([_u64(_addr(d_cw) + 8 * cw_stride * i, N) for i in range(3)])
        r = orc.fri_fold(cw, [int(alpha[i]) for i in range(3)], offset, omega)
        for i in range(3):
            _u64(_addr(d_next) + 8 * next_stride * i, N // 2)[:] = r[i]
        if _addr(d_next_nodes):
            _u8(_addr(d_next_nodes), 64 * N)[:] = orc.merkle_field(_otpl(tpl), r).reshape(-1)
        return 0

    def b2s_gather(self, d_planes, stride, q, h_idx, n_idx, h_out, stream):
        self.launches += 1
        idx = _u64(_addr(h_idx), n_idx)
        out = _u64(_addr(h_out), n_idx * q).reshape(n_idx, q)
        for pl in range(q):
            col = _u64(_addr(d_planes) + 8 * stride * pl, int(idx.max()) + 1)
            out[:, pl] = col[idx.astype(np.int64)]
        return 0


    def b2s_quotients(self, d_cw, N, width, shift, nc, h_off, h_coeffs, h_factors, max_factors, kind, height, oinv,
                      offset, omega, d_out, h_flag, h_base_columns, stream):
        self.launches += 1
        cw = _u64(_addr(d_cw), width * 3 * N).reshape(width, 3, N)
        if _addr(h_base_columns):  # the caller's claim must hold, or the device would silently drop planes 1 and 2
            flags = _u8(_addr(h_base_columns), width)
            assert not cw[flags != 0, 1:, :].any(), "a column flagged base-field has non-zero upper planes"
        off = np.ctypeslib.as_array((C.c_uint32 * (nc + 1)).from_address(_addr(h_off)))
        m = int(off[nc])
        coeffs = _u64(_addr(h_coeffs), 3 * m).reshape(m, 3) if m else np.zeros((0, 3), dtype=np.uint64)
        fac = (np.ctypeslib.as_array((C.c_uint32 * (m * max_factors)).from_address(_addr(h_factors)))
               .reshape(m, max_factors) if m * max_factors else np.zeros((m, 0), dtype=np.uint32))
        out, flag = orc.quotients(cw, shift, off, coeffs, fac, kind, height, oinv, offset, omega)
        if nc:
            _u64(_addr(d_out), nc * 3 * N).reshape(nc, 3, N)[:] = out
        if _addr(h_flag):
            _set_i32(h_flag, flag)
        else:  # the caller claimed no zerofier can vanish on this domain and skipped the read-back: hold it to that
            assert not flag, "b2s_quotients called without a zero flag, but a zerofier vanishes on the domain"
        return 0

    def b2s_open_multi(self, h_planes, h_strides, n_planes, h_nodes, h_npo2, h_counts, h_indices, n_sets, h_values,
                       h_paths, stream):
        self.launches += 1
        planes, strides = _u64(_addr(h_planes), n_sets), _u64(_addr(h_strides), n_sets)
        nodes, npo2 = _u64(_addr(h_nodes), n_sets), _u64(_addr(h_npo2), n_sets)
        counts = np.ctypeslib.as_array((C.c_uint32 * n_sets).from_address(_addr(h_counts)))
        total = int(counts.sum())
        idx = _u64(_addr(h_indices), total)
        values = _u64(_addr(h_values), total * n_planes).reshape(total, n_planes)
        pos = ppos = 0
        for s in range(n_sets):
            n = int(npo2[s])
            depth = n.bit_length() - 1 if nodes[s] else 0
            for q in range(int(counts[s])):
                i = int(idx[pos])
                if planes[s]:
                    for pl in range(n_planes):
                        values[pos, pl] = _u64(int(planes[s]) + 8 * (int(strides[s]) * pl + i), 1)[0]
                if nodes[s]:
                    assert i < n
                    heap = _u8(int(nodes[s]), 128 * n).reshape(2 * n, 64)
                    out = _u8(_addr(h_paths) + ppos, depth * 64).reshape(depth, 64) if depth else None
                    for lvl in range(depth):
                        out[lvl] = heap[((n + i) >> lvl) ^ 1]
                    ppos += depth * 64
                pos += 1
        return 0

    def b2s_combination(self, h_cols, h_strides, h_planes, h_wa, h_wb, h_shifts, n_cols, N, offset, omega, d_out,
                        out_stride, stream):
        self.launches += 1
        ptrs = _u64(_addr(h_cols), n_cols)
        strides = _u64(_addr(h_strides), n_cols)
        planes = np.ctypeslib.as_array((C.c_uint32 * n_cols).from_address(_addr(h_planes)))
        cols = [np.stack([_u64(int(ptrs[c]) + 8 * int(strides[c]) * q, N) for q in range(int(planes[c]))])
                for c in range(n_cols)]
        out = orc.combination(cols, _u64(_addr(h_wa), 3 * n_cols), _u64(_addr(h_wb), 3 * n_cols),
                              _u64(_addr(h_shifts), n_cols), offset, omega)
        for q in range(3):
            _u64(_addr(d_out) + 8 * out_stride * q, N)[:] = out[q]
        return 0

    def b2s_dist_twiddle_transpose(self, d_in, in_stride, rows, cols, row_base, omega, tw_mul, out_ptrs, n_peers,
                                   out_row_stride, out_col_offset, stream):
        self.launches += 1
        P = orc.P
        cpp = cols // n_peers
        for r in range(rows):
            row = _u64(_addr(d_in) + 8 * in_stride * r, cols)
            j1 = row_base + r
            for c in range(cols):
                v = int(row[c]) * pow(omega, tw_mul * j1 * c, P) % P
                dst = _u64(_addr(out_ptrs[c // cpp]) + 8 * ((c % cpp) * out_row_stride + out_col_offset + r), 1)
                dst[0] = v
        return 0

    def b2s_block_permute(self, d_in, d_out, A, B, Cn, stream):
        self.launches += 1
        src = _u64(_addr(d_in), A * B * Cn).reshape(A, B, Cn)
        _u64(_addr(d_out), A * B * Cn).reshape(B, A, Cn)[:] = src.transpose(1, 0, 2)
        return 0


def fake_engine():
    from stark_brainfuck_b200 import Engine
    return Engine(lib=FakeLib(), device="cpu")
