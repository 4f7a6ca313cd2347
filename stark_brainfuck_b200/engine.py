"""Array-level host API over the C ABI: device-resident planes in, device-resident planes out.

torch is used only for device memory (int64 / uint8 tensors as byte buffers) and streams;
all arithmetic happens inside libb2s.so.  Shapes: a base-field vector is (1, n) int64, an
extension-field vector (3, n) int64 (planes c0, c1, c2), a batch of columns (q, n).  The
int64 dtype is a container for uint64 bit patterns.

`Engine(lib=..., device=...)` exists so that tests/ can inject a host-memory backend with
the same C ABI; the default constructor loads libb2s.so, initialises the CUDA device and
raises if either is missing -- there is no CPU fallback in the product path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

P = 18446744069414584321


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device_index=0, lib=None, device=None):
        if lib is None:
            if not torch.cuda.is_available():
                raise _lib.B2SError("no CUDA device: stark_brainfuck_b200 has no CPU fallback")
            lib = _lib.load()
            torch.cuda.set_device(device_index)
            _lib.check(lib, lib.b2s_init(device_index))
            device = torch.device("cuda", device_index)
        self.lib = lib
        self.device = torch.device(device)
        self.templates = {}  # name -> LeafTemplates

    # ---- plumbing ----------------------------------------------------------------
    def stream_ptr(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    # Device-memory primitives.  Everything the glue does to device memory outside the C ABI goes through
    # these methods (allocation, fill, host -> device, device -> host, device -> device), so that a recording
    # engine (tests/trace_backend.py) sees the complete command stream of a proof.
    def alloc(self, shape, dtype=torch.int64, zero=False):
        return (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)

    def upload_into(self, dst, arr):
        """host numpy array -> the device view `dst` (same shape; uint64 data into int64 views)"""
        a = np.ascontiguousarray(arr)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        if not a.flags.writeable:  # e.g. np.frombuffer over bytes: torch refuses to wrap read-only memory
            a = a.copy()
        dst.copy_(torch.from_numpy(a).reshape(dst.shape))
        return dst

    def zero(self, dst):
        dst.zero_()
        return dst

    def copy(self, dst, src):
        """device -> device"""
        dst.copy_(src)
        return dst

    def upload(self, arr, pinned=False):
        """numpy uint64 array (n,) or (q, n) -> device int64 tensor (q, n)"""
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        if pinned and self.device.type == "cuda":
            t = torch.from_numpy(a.view(np.int64)).pin_memory()
            return t.to(self.device, non_blocking=True)
        return self.upload_into(self.alloc(a.shape), a)

    def upload_bytes(self, data):
        """bytes / uint8 array -> device uint8 tensor of the same shape"""
        a = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, dtype=np.uint8)
        return self.upload_into(self.alloc(a.shape, torch.uint8), a)

    def download(self, t):
        """device tensor -> numpy uint64 array with the same shape (synchronises)"""
        return t.detach().cpu().contiguous().numpy().view(np.uint64)

    def empty(self, planes, n):
        return self.alloc((planes, n))

    def zeros(self, planes, n):
        return self.alloc((planes, n), zero=True)

    def check(self, rc):
        _lib.check(self.lib, rc)

    def launch_count(self):
        return int(self.lib.b2s_launch_count())

    # ---- NTT family ---------------------------------------------------------------
    def ntt(self, x, log_n, omega, offset=1, inverse=False, out=None):
        """code/ntt.py:4-42, :164-174 on planes.  x: (q, n_in) with n_in <= 2^log_n
        (forward: missing coefficients are zero padding; inverse: n_in == n)."""
        q, n_in = x.shape
        n = 1 << log_n
        assert x.dtype == torch.int64 and x.stride(1) == 1
        if out is None:
            out = self.empty(q, n)
        self.check(self.lib.b2s_ntt(_ptr(x), x.stride(0) if q > 1 else max(n_in, 1), n_in, _ptr(out),
                                    out.stride(0) if q > 1 else n, log_n, q, omega, offset, 1 if inverse else 0,
                                    self.stream_ptr()))
        return out

    def ntt_host(self, h_in, log_n, omega, offset=1, inverse=False, h_out=None):
        """same through host buffers (pinned int64 CPU tensors): H2D + kernels + D2H"""
        q, n_in = h_in.shape
        n = 1 << log_n
        if h_out is None:
            h_out = torch.empty((q, n), dtype=torch.int64, pin_memory=self.device.type == "cuda")
        self.check(self.lib.b2s_ntt_host(_ptr(h_in), h_in.stride(0), n_in, _ptr(h_out), h_out.stride(0), log_n, q,
                                         omega, offset, 1 if inverse else 0))
        return h_out

    def ntt_timed(self, x, log_n, omega, offset=1, inverse=False, out=None, iters=10):
        q, n_in = x.shape
        n = 1 << log_n
        if out is None:
            out = self.empty(q, n)
        ms = C.c_float(0)
        self.check(self.lib.b2s_ntt_timed(_ptr(x), x.stride(0) if q > 1 else max(n_in, 1), n_in, _ptr(out),
                                          out.stride(0) if q > 1 else n, log_n, q, omega, offset,
                                          1 if inverse else 0, self.stream_ptr(), iters, C.byref(ms)))
        return ms.value, out

    def scale(self, x, factor):
        """code/univariate.py:168-169.  factor: int (base field) or 3 ints (extension field)"""
        q, n = x.shape
        f = (C.c_uint64 * 3)(*([factor, 0, 0] if isinstance(factor, int) else list(factor)))
        out = self.alloc(x.shape)
        self.check(self.lib.b2s_scale(_ptr(x), x.stride(0), _ptr(out), out.stride(0), n, q, f, self.stream_ptr()))
        return out

    def eval_points(self, coeffs, points):
        """code/univariate.py:145-154 on arbitrary points"""
        cq, m = coeffs.shape
        pq, k = points.shape
        out = self.empty(max(cq, pq), k)
        self.check(self.lib.b2s_eval_points(_ptr(coeffs), coeffs.stride(0), cq, m, _ptr(points), points.stride(0), pq,
                                            k, _ptr(out), out.stride(0), self.stream_ptr()))
        return out

    # ---- Merkle -------------------------------------------------------------------
    def merkle_field(self, planes, tpl):
        """code/merkle.py:8-41 over field-element leaves; returns nodes (2n, 64) uint8"""
        q, n = planes.shape
        assert q == tpl.n_slots
        nodes = self.alloc((2 * n, 64), torch.uint8)
        self.check(self.lib.b2s_merkle_field(_ptr(planes), planes.stride(0), n, C.byref(tpl), _ptr(nodes),
                                             self.stream_ptr()))
        return nodes

    def merkle_blobs(self, blobs):
        """arbitrary leaves already pickled by the caller (code/test_merkle.py:57-61)"""
        n = len(blobs)
        npo2 = 1
        while npo2 < n:
            npo2 *= 2
        offs = np.zeros(n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in blobs])
        d_data = self.upload_bytes(b"".join(blobs) + b"\0" * 8)
        d_offs = self.upload(offs)
        nodes = self.alloc((2 * npo2, 64), torch.uint8)
        self.check(self.lib.b2s_merkle_blobs(_ptr(d_data), _ptr(d_offs), n, npo2, _ptr(nodes), self.stream_ptr()))
        return nodes

    def merkle_rows(self, planes, modes, tpl, seg_off, n, salts=None, salt_pre=b"", salt_suf=b"", rows=None,
                    nodes=None, build_upper=True):
        """code/salted_merkle.py:25-35 over rows of codewords.  planes: list of (n,) device int64 views in row order;
        modes / tpl / seg_off: the row template (marshal.RowTemplate); salts: (n, salt_len) uint8 device tensor.
        rows: optional device int32 tensor of row indices.  Returns (nodes (2n, 64) uint8, exception rows)."""
        for t in planes:
            assert t.dtype == torch.int64 and t.shape == (n,) and t.stride(0) == 1
        ptrs = np.array([t.data_ptr() for t in planes], dtype=np.uint64)
        modes = np.ascontiguousarray(modes, dtype=np.uint8)
        seg = np.ascontiguousarray(seg_off, dtype=np.uint32)
        tpl = np.frombuffer(bytes(tpl) + b"\0", dtype=np.uint8)
        pre = np.frombuffer(bytes(salt_pre) + b"\0", dtype=np.uint8)
        suf = np.frombuffer(bytes(salt_suf) + b"\0", dtype=np.uint8)
        if nodes is None:
            nodes = self.alloc((2 * n, 64), torch.uint8)
        count = n if rows is None else rows.numel()
        exc = np.zeros(max(count, 1), dtype=np.uint32)
        n_exc = C.c_uint32(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        if salts is not None:
            assert salts.dtype == torch.uint8 and salts.is_contiguous() and salts.shape[0] == n
        self.check(self.lib.b2s_merkle_rows(vp(ptrs), vp(modes), len(planes), n, vp(tpl), vp(seg), len(seg) - 2,
                                            _ptr(salts) if salts is not None else None,
                                            salts.shape[1] if salts is not None else 0, vp(pre), len(salt_pre), vp(suf),
                                            len(salt_suf), _ptr(rows) if rows is not None else None,
                                            0 if rows is None else count, _ptr(nodes), 1 if build_upper else 0, vp(exc),
                                            C.byref(n_exc), self.stream_ptr()))
        return nodes, exc[:n_exc.value].copy()

    def merkle_upper(self, nodes):
        """inner nodes above the digests in slots [npo2, 2 npo2) of `nodes` ((2 npo2, 64) uint8), in place"""
        self.check(self.lib.b2s_merkle_upper(_ptr(nodes), nodes.shape[0] // 2, self.stream_ptr()))
        return nodes

    def merkle_open(self, nodes, indices):
        """code/merkle.py:46-52 for several indices: list of lists of 64-byte digests"""
        npo2 = nodes.shape[0] // 2
        depth = npo2.bit_length() - 1
        if depth == 0 or not indices:
            return [[] for _ in indices]
        idx = np.asarray(indices, dtype=np.uint64)
        out = np.empty((len(indices), depth, 64), dtype=np.uint8)
        self.check(self.lib.b2s_merkle_open(_ptr(nodes), npo2, idx.ctypes.data_as(C.c_void_p), len(indices),
                                            out.ctypes.data_as(C.c_void_p), self.stream_ptr()))
        return [[out[q, j].tobytes() for j in range(depth)] for q in range(len(indices))]

    def open_multi(self, sets):
        """code/fri.py:141-176 in one device call.  sets: [(planes (q, n) or None, nodes (2 npo2, 64) or None, indices)];
        returns [(values (len, q) uint64 or None, paths (len, depth, 64) uint8 or None)] in the same order."""
        sets = [(p_, n_, list(i_)) for p_, n_, i_ in sets]
        ns = len(sets)
        total = sum(len(i_) for _, _, i_ in sets)
        if total == 0:
            return [(None, None)] * ns
        q = max([p_.shape[0] for p_, _, _ in sets if p_ is not None] + [1])
        planes = np.array([p_.data_ptr() if p_ is not None else 0 for p_, _, _ in sets], dtype=np.uint64)
        strides = np.array([(p_.stride(0) if p_.shape[0] > 1 else p_.shape[1]) if p_ is not None else 0
                            for p_, _, _ in sets], dtype=np.uint64)
        nodes = np.array([n_.data_ptr() if n_ is not None else 0 for _, n_, _ in sets], dtype=np.uint64)
        npo2 = np.array([n_.shape[0] // 2 if n_ is not None else 0 for _, n_, _ in sets], dtype=np.uint64)
        counts = np.array([len(i_) for _, _, i_ in sets], dtype=np.uint32)
        idx = np.array([i for _, _, i_ in sets for i in i_], dtype=np.uint64)
        depths = [int(m).bit_length() - 1 if m else 0 for m in npo2]
        values = np.zeros((total, q), dtype=np.uint64)
        paths = np.zeros(sum(c * d for c, d in zip(counts.tolist(), depths)) * 64 + 1, dtype=np.uint8)
        for p_, _, _ in sets:
            assert p_ is None or (p_.shape[0] == q and (p_.shape[1] == 1 or p_.stride(1) == 1))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.check(self.lib.b2s_open_multi(vp(planes), vp(strides), q, vp(nodes), vp(npo2), vp(counts), vp(idx), ns,
                                           vp(values), vp(paths), self.stream_ptr()))
        out, vpos, ppos = [], 0, 0
        for (p_, n_, i_), d in zip(sets, depths):
            c = len(i_)
            v = values[vpos:vpos + c] if p_ is not None else None
            pa = paths[ppos:ppos + c * d * 64].reshape(c, d, 64) if n_ is not None else None
            vpos += c
            ppos += c * d * 64 if n_ is not None else 0
            out.append((v, pa))
        return out

    def root(self, nodes):
        return bytes(self.download_bytes(nodes[1]))

    def download_bytes(self, t):
        return t.detach().cpu().contiguous().numpy().tobytes()

    # ---- FRI ----------------------------------------------------------------------
    def fri_fold(self, cw, alpha, offset, omega, tpl=None):
        """code/fri.py:127-128 (+ Merkle tree of the folded codeword when tpl is given)"""
        q, N = cw.shape
        assert q == 3
        nxt = self.empty(3, N // 2)
        nodes = self.alloc((N, 64), torch.uint8) if tpl is not None else None
        a = (C.c_uint64 * 3)(*[int(v) for v in alpha])
        self.check(self.lib.b2s_fri_fold(_ptr(cw), cw.stride(0), N, a, offset, omega, _ptr(nxt), nxt.stride(0),
                                         C.byref(tpl) if tpl is not None else None,
                                         _ptr(nodes) if nodes is not None else None, self.stream_ptr()))
        return nxt, nodes

    # ---- quotient codewords ---------------------------------
    def quotients(self, cw, shift, mono_off, coeffs, factors, kind, height, omicron_inv, offset, omega,
                  base_columns=None, check_zerofier=True):
        """code/table.py:155-286 for one table.  cw: (width, 3, N) int64 device tensor; the constraint
        program as numpy arrays (mono_off (C+1,) uint32, coeffs (M, 3) uint64, factors (M, F) uint32).
        base_columns: optional per-codeword flags, True where the caller knows planes 1 and 2 are zero (a lifted
        base-field column); None lets the library scan.
        check_zerofier=False: the caller has ruled out a zerofier vanishing on the domain; the flag is not read
        back and the call does not synchronise.
        Returns ((C, 3, N) device tensor, True if a zerofier vanishes on the domain)."""
        width, three, N = cw.shape
        assert three == 3 and cw.is_contiguous()
        mono_off = np.ascontiguousarray(mono_off, dtype=np.uint32)
        nc = len(mono_off) - 1
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 3)
        factors = np.ascontiguousarray(factors, dtype=np.uint32).reshape(len(coeffs), -1)
        out = self.alloc((nc, 3, N))
        flag = C.c_int(0)
        base = None if base_columns is None else np.ascontiguousarray(base_columns, dtype=np.uint8).reshape(width)
        self.check(self.lib.b2s_quotients(_ptr(cw), N, width, shift, nc, mono_off.ctypes.data_as(C.c_void_p),
                                          coeffs.ctypes.data_as(C.c_void_p), factors.ctypes.data_as(C.c_void_p),
                                          factors.shape[1] if factors.size else 0, kind, height, omicron_inv, offset,
                                          omega, _ptr(out), C.byref(flag) if check_zerofier else None,
                                          None if base is None else base.ctypes.data_as(C.c_void_p), self.stream_ptr()))
        return out, bool(flag.value)

    # ---- nonlinear combination --------------------------------
    def combination(self, columns, wa, wb, shifts, N, offset, omega):
        """code/brainfuck_stark.py:241-298: sum_c (wa_c + wb_c x^shift_c) col_c over the domain offset*omega^j.
        columns: device tensors (1, N) (base field) or (3, N); wa, wb: (n_cols, 3) uint64; shifts: (n_cols,).
        Returns (3, N) planes."""
        n = len(columns)
        ptrs = np.array([t.data_ptr() for t in columns], dtype=np.uint64)
        strides = np.array([t.stride(0) if t.shape[0] > 1 else N for t in columns], dtype=np.uint64)
        planes = np.array([t.shape[0] for t in columns], dtype=np.uint32)
        for t in columns:
            assert t.dtype == torch.int64 and t.shape[1] == N and t.stride(1) == 1
        wa = np.ascontiguousarray(wa, dtype=np.uint64).reshape(n, 3)
        wb = np.ascontiguousarray(wb, dtype=np.uint64).reshape(n, 3)
        shifts = np.ascontiguousarray(shifts, dtype=np.uint64).reshape(n)
        out = self.empty(3, N)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.check(self.lib.b2s_combination(vp(ptrs), vp(strides), vp(planes), vp(wa), vp(wb), vp(shifts), n, N, offset,
                                            omega, _ptr(out), out.stride(0), self.stream_ptr()))
        return out

    def gather(self, planes, indices):
        """planes[:, indices] to the host as numpy (len(indices), q) uint64"""
        q, n = planes.shape
        if not len(indices):
            return np.empty((0, q), dtype=np.uint64)
        idx = np.asarray(indices, dtype=np.uint64)
        out = np.empty((len(indices), q), dtype=np.uint64)
        self.check(self.lib.b2s_gather(_ptr(planes), planes.stride(0), q, idx.ctypes.data_as(C.c_void_p), len(indices),
                                       out.ctypes.data_as(C.c_void_p), self.stream_ptr()))
        return out


_default = None


def default_engine():
    """process-wide engine on cuda:LOCAL_RANK (created on first use; raises without CUDA)"""
    global _default
    if _default is None:
        import os
        _default = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _default


def set_default_engine(e):
    global _default
    _default = e
