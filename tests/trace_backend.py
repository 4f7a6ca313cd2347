"""TEST-ONLY: record the complete device command stream of a computation (every allocation, host<->device copy
and C-ABI call of include/b2s.h, with the content of every host input and a digest of every output) and replay
it against the real library on a GPU.

Why: BrainfuckStark.prove() of the unmodified reference can only run where the reference checkout is -- the
authoring container, which has no GPU.  There, the drop-in runs over the host-memory test backend
(tests/fake_backend.py = the CPU oracle) and this module records what the glue asks of the engine: a trace that
needs neither the reference nor the glue to be replayed.  `pytest -m gpu` replays it through libb2s.so on the
B200 and compares every output -- kernel by kernel and every byte that went back to the host (Merkle roots,
openings, codewords), which together determine the proof.

Trace file: zlib( u64 header length | JSON header | concatenated literal blobs ).  Host inputs that are slices of
the seeded `urandom` stream (the salts of the salted trees) are stored as (offset, length) into that stream.
"""
import bisect
import ctypes as C
import hashlib
import json
import random
import struct
import zlib

import numpy as np
import torch

from stark_brainfuck_b200 import _lib
from stark_brainfuck_b200.engine import Engine


class SeededUrandom:
    """os.urandom stand-in of tests/e2e_prove_dropin.py and tests/golden/make_golden.py:
    bytes(R.getrandbits(8) for _ in range(n)) with R = random.Random(seed) -- the top byte of consecutive
    32-bit Mersenne-Twister outputs -- generated in bulk.  Keeps the whole stream (the recorder refers to it)."""

    def __init__(self, seed, chunk=1 << 16):
        self.R = random.Random(seed)
        self.stream = bytearray()
        self.pos = 0
        self.chunk = chunk

    def _more(self, k):
        words = np.frombuffer(self.R.getrandbits(32 * k).to_bytes(4 * k, "little"), dtype="<u4")
        self.stream += (words >> 24).astype(np.uint8).tobytes()

    def __call__(self, n):
        while self.pos + n > len(self.stream):
            self._more(max(self.chunk, n))
        out = bytes(self.stream[self.pos:self.pos + n])
        self.pos += n
        return out


def _sha(b):
    return hashlib.sha256(b).hexdigest()[:32]


def _address(a):
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, C.c_void_p):
        return a.value or 0
    if hasattr(a, "_obj"):  # byref(x)
        return C.addressof(a._obj)
    if isinstance(a, (C.Array, C.Structure)):
        return C.addressof(a)
    if hasattr(a, "contents"):
        return C.addressof(a.contents)
    raise TypeError("cannot take the address of %r" % (a,))


def _ilog2(n):
    return int(n).bit_length() - 1


# ---- what every entry point reads and writes -----------------------------------------------------------------
# kinds: s scalar | d device pointer | S stream | T leaf templates (nullable) | A3 three u64 in host memory |
#        ("hi", nbytes(a)) host input | ("ho", nbytes(a)) host output | ("hp", count(a)) host table of device pointers
# dev_out(a) -> [(arg index, byte offset, rows, row bytes, row stride bytes)] regions a call writes
def _q_monos(a):
    off = np.frombuffer(C.string_at(_address(a[5]), 4 * (a[4] + 1)), dtype=np.uint32)
    return int(off[a[4]])


def _open_counts(a):
    return np.frombuffer(C.string_at(_address(a[5]), 4 * a[7]), dtype=np.uint32)


def _open_path_bytes(a):
    npo2 = np.frombuffer(C.string_at(_address(a[4]), 8 * a[7]), dtype=np.uint64)
    nodes = np.frombuffer(C.string_at(_address(a[3]), 8 * a[7]), dtype=np.uint64)
    return int(sum(int(c) * _ilog2(m) * 64 for c, m, nd in zip(_open_counts(a), npo2, nodes) if nd)) + 1


def _rows_count(a):
    return a[14] if _address(a[13]) else a[3]


SPECS = {
    "b2s_ntt": (["d", "s", "s", "d", "s", "s", "s", "s", "s", "s", "S"],
                lambda a: [(3, 0, a[6], 8 << a[5], 8 * a[4])]),
    "b2s_scale": (["d", "s", "d", "s", "s", "s", "A3", "S"], lambda a: [(2, 0, a[5], 8 * a[4], 8 * a[3])]),
    "b2s_eval_points": (["d", "s", "s", "s", "d", "s", "s", "s", "d", "s", "S"],
                        lambda a: [(8, 0, max(a[2], a[6]), 8 * a[7], 8 * a[9])]),
    "b2s_merkle_field": (["d", "s", "s", "T", "d", "S"], lambda a: [(4, 64, 1, 128 * a[2] - 64, 0)]),
    "b2s_merkle_blobs": (["d", "d", "s", "s", "d", "S"], lambda a: [(4, 64, 1, 128 * a[3] - 64, 0)]),
    "b2s_merkle_rows": (["hp2", ("hi", lambda a: a[2]), "s", "s", ("hi", lambda a: int(np.frombuffer(C.string_at(
        _address(a[5]), 4 * (a[6] + 2)), dtype=np.uint32)[-1])), ("hi", lambda a: 4 * (a[6] + 2)), "s", "d", "s",
        ("hi", lambda a: a[10]), "s", ("hi", lambda a: a[12]), "s", "d", "s", "d", "s",
        ("ho", lambda a: 4 * _rows_count(a)), ("ho", lambda a: 4), "S"],
        # (a row list or rows of another shape leave other leaf slots untouched: nothing stable to compare)
        lambda a: [] if _address(a[13]) or np.frombuffer(C.string_at(_address(a[18]), 4), dtype=np.uint32)[0]
        else [(15, 64 * a[3], 1, 64 * a[3], 0)]),
    "b2s_merkle_upper": (["d", "s", "S"], lambda a: [(0, 64, 1, 128 * a[1] - 64, 0)]),
    "b2s_merkle_open": (["d", "s", ("hi", lambda a: 8 * a[3]), "s", ("ho", lambda a: a[3] * _ilog2(a[1]) * 64), "S"],
                        lambda a: []),
    "b2s_fri_fold": (["d", "s", "s", "A3", "s", "s", "d", "s", "T", "d", "S"],
                     lambda a: [(6, 0, 3, 8 * (a[2] // 2), 8 * a[7])] +
                               ([(9, 64, 1, 64 * a[2] - 64, 0)] if _address(a[9]) else [])),
    "b2s_gather": (["d", "s", "s", ("hi", lambda a: 8 * a[4]), "s", ("ho", lambda a: 8 * a[4] * a[2]), "S"],
                   lambda a: []),
    "b2s_quotients": (["d", "s", "s", "s", "s", ("hi", lambda a: 4 * (a[4] + 1)), ("hi", lambda a: 24 * _q_monos(a)),
                       ("hi", lambda a: 4 * _q_monos(a) * a[8]), "s", "s", "s", "s", "s", "s", "d", ("ho", lambda a: 4),
                       ("hi", lambda a: a[2]), "S"], lambda a: [(14, 0, 3 * a[4], 8 * a[1], 8 * a[1])]),
    "b2s_open_multi": (["hp7", ("hi", lambda a: 8 * a[7]), "s", "hp7", ("hi", lambda a: 8 * a[7]),
                        ("hi", lambda a: 4 * a[7]), ("hi", lambda a: 8 * int(_open_counts(a).sum())), "s",
                        ("ho", lambda a: 8 * int(_open_counts(a).sum()) * a[2]), ("ho", _open_path_bytes), "S"],
                       lambda a: []),
    "b2s_combination": (["hp6", ("hi", lambda a: 8 * a[6]), ("hi", lambda a: 4 * a[6]), ("hi", lambda a: 24 * a[6]),
                         ("hi", lambda a: 24 * a[6]), ("hi", lambda a: 8 * a[6]), "s", "s", "s", "s", "d", "s", "S"],
                        lambda a: [(10, 0, 3, 8 * a[7], 8 * a[11])]),
}


class Recorder:
    def __init__(self, urandom=None):
        self.events = []
        self.blobs = []
        self.blob_pos = {}
        self.size = 0
        self.urandom = urandom
        self.starts, self.ranges = [], []  # sorted allocation table: start addresses, (start, end, id)
        self.n_allocs = 0

    # -- address space
    def register(self, t, zero):
        base, nbytes = t.data_ptr(), t.numel() * t.element_size()
        if nbytes == 0:
            self.events.append({"op": "alloc", "id": self.n_allocs, "nbytes": 0, "zero": False})
            self.n_allocs += 1
            return self.n_allocs - 1
        i = bisect.bisect_left(self.starts, base)
        while i > 0 and self.ranges[i - 1][1] > base:  # stale entries of freed tensors that overlap
            i -= 1
        j = i
        while j < len(self.ranges) and self.ranges[j][0] < base + max(nbytes, 1):
            j += 1
        del self.starts[i:j], self.ranges[i:j]
        k = self.n_allocs
        self.n_allocs += 1
        self.starts.insert(i, base)
        self.ranges.insert(i, (base, base + max(nbytes, 1), k))
        self.events.append({"op": "alloc", "id": k, "nbytes": nbytes, "zero": bool(zero)})
        return k

    def locate(self, addr):
        if addr == 0:
            return None
        i = bisect.bisect_right(self.starts, addr) - 1
        if i < 0 or not self.ranges[i][0] <= addr < self.ranges[i][1]:
            raise KeyError("address %#x is not inside a recorded allocation" % addr)
        return [self.ranges[i][2], addr - self.ranges[i][0]]

    def view(self, t):
        return {"at": self.locate(t.data_ptr()) if t.numel() else None, "shape": list(t.shape),
                "strides": [s * t.element_size() for s in t.stride()], "item": t.element_size()}

    # -- literal data
    def blob(self, data):
        data = bytes(data)
        if self.urandom is not None and len(data) >= 1024:
            at = self.urandom.stream.find(data)
            if at >= 0:
                return {"rng": [at, len(data)]}
        key = hashlib.sha256(data).digest()
        if key not in self.blob_pos:
            self.blob_pos[key] = (self.size, len(data))
            self.blobs.append(data)
            self.size += len(data)
        return {"blob": list(self.blob_pos[key])}

    def save(self, path, meta):
        header = json.dumps({"meta": meta, "events": self.events,
                             "rng": ({"seed_stream_bytes": len(self.urandom.stream)} if self.urandom else None)}).encode()
        with open(path, "wb") as f:
            f.write(zlib.compress(struct.pack("<Q", len(header)) + header + b"".join(self.blobs), 9))


class TracingLib:
    """forwards every C-ABI call to `inner` and records it"""

    def __init__(self, inner, rec):
        self._inner, self._rec = inner, rec

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if name not in SPECS:
            if name.startswith("b2s_") and name not in ("b2s_last_error", "b2s_launch_count"):
                raise NotImplementedError("tests/trace_backend.py has no recording rule for %s" % name)
            return fn
        kinds, dev_out = SPECS[name]
        rec = self._rec

        def call(*args):
            a = scal = list(args)  # the size expressions read scalars and host arrays
            enc = []
            for i, k in enumerate(kinds):
                x = a[i]
                if k == "s":
                    enc.append(int(x))
                elif k == "S":
                    enc.append("stream")
                elif k == "d":
                    enc.append({"p": rec.locate(_address(x))})
                elif k == "T":
                    enc.append({"tpl": rec.blob(C.string_at(_address(x), C.sizeof(_lib.LeafTemplates)))} if _address(x)
                               else None)
                elif k == "A3":
                    enc.append({"a3": [int(v) for v in np.frombuffer(C.string_at(_address(x), 24), dtype=np.uint64)]})
                elif isinstance(k, str) and k.startswith("hp"):
                    cnt = a[int(k[2:])]
                    ptrs = np.frombuffer(C.string_at(_address(x), 8 * cnt), dtype=np.uint64)
                    enc.append({"ptrs": [rec.locate(int(p)) for p in ptrs]})
                elif k[0] == "hi":
                    nb = k[1](scal)
                    enc.append({"h": rec.blob(C.string_at(_address(x), nb)) if nb and _address(x) else None,
                                "null": _address(x) == 0})
                elif k[0] == "ho":  # a NULL output pointer (b2s_quotients' optional flag) is recorded as such
                    enc.append({"o": k[1](scal), "null": True} if _address(x) == 0 else {"o": k[1](scal)})
            rc = fn(*args)
            outs = []
            for i, k in enumerate(kinds):
                if not isinstance(k, str) and k[0] == "ho" and _address(a[i]):
                    outs.append({"arg": i, "sha": _sha(C.string_at(_address(a[i]), k[1](scal)))})
            devs = []
            if rc == 0:
                for arg, off, rows, row_bytes, stride in dev_out(scal):
                    base = _address(a[arg]) + off
                    data = b"".join(C.string_at(base + r * stride, row_bytes) for r in range(rows))
                    devs.append({"arg": arg, "off": off, "rows": rows, "row_bytes": row_bytes, "stride": stride,
                                 "sha": _sha(data)})
            rec.events.append({"op": "call", "fn": name, "args": enc, "rc": int(rc), "host_out": outs, "dev_out": devs})
            return rc
        return call


class TracingEngine(Engine):
    """Engine over the host-memory backend that records everything it does to `device` memory"""

    def __init__(self, inner_lib, rec):
        super().__init__(lib=TracingLib(inner_lib, rec), device="cpu")
        self.rec = rec

    def alloc(self, shape, dtype=torch.int64, zero=False):
        t = super().alloc(shape, dtype, zero)
        self.rec.register(t, zero)
        return t

    def upload_into(self, dst, arr):
        a = np.ascontiguousarray(arr)
        self.rec.events.append({"op": "h2d", "dst": self.rec.view(dst), "data": self.rec.blob(a.tobytes())})
        return super().upload_into(dst, arr)

    def zero(self, dst):
        self.rec.events.append({"op": "zero", "dst": self.rec.view(dst)})
        return super().zero(dst)

    def copy(self, dst, src):
        self.rec.events.append({"op": "d2d", "dst": self.rec.view(dst), "src": self.rec.view(src)})
        return super().copy(dst, src)

    def download(self, t):
        out = super().download(t)
        self.rec.events.append({"op": "d2h", "src": self.rec.view(t), "sha": _sha(out.tobytes())})
        return out

    def download_bytes(self, t):
        out = super().download_bytes(t)
        self.rec.events.append({"op": "d2h", "src": self.rec.view(t), "sha": _sha(out)})
        return out


# ---- replay ------------------------------------------------------------------------------------------------
def load_trace(path):
    with open(path, "rb") as f:
        raw = zlib.decompress(f.read())
    (hl,) = struct.unpack("<Q", raw[:8])
    header = json.loads(raw[8:8 + hl])
    return header, memoryview(raw)[8 + hl:]


def replay(path, engine, urandom_seed=1234, check_kernels=True, check_reads=True):
    """Run the recorded command stream on `engine` (a real Engine on a CUDA device).  Every device -> host read and
    (check_kernels) every region a call wrote is compared with the recording.  Returns a summary dict."""
    header, blobs = load_trace(path)
    lib, dev = engine.lib, engine.device
    stream = None
    if header.get("rng"):
        u = SeededUrandom(urandom_seed)
        u(header["rng"]["seed_stream_bytes"])
        stream = bytes(u.stream)

    def data_of(ref):
        if ref is None:
            return b""
        if "rng" in ref:
            at, n = ref["rng"]
            return stream[at:at + n]
        at, n = ref["blob"]
        return bytes(blobs[at:at + n])

    mem = {}

    def addr(loc):
        return 0 if loc is None else mem[loc[0]].data_ptr() + loc[1]

    def view(v):
        if v["at"] is None:
            return torch.empty(v["shape"], dtype=torch.uint8, device=dev)
        k, off = v["at"]
        item = v["item"]
        flat = mem[k]  # uint8
        dt = {1: torch.uint8, 4: torch.int32, 8: torch.int64}[item]
        assert off % item == 0
        return torch.as_strided(flat.view(dt), v["shape"], [s // item for s in v["strides"]], off // item)

    n_calls = n_checked = n_d2h = 0
    keep = []
    for ev in header["events"]:
        op = ev["op"]
        if op == "alloc":
            nb = (ev["nbytes"] + 15) // 16 * 16 + 16
            mem[ev["id"]] = (torch.zeros if ev["zero"] else torch.empty)(nb, dtype=torch.uint8, device=dev)
        elif op == "h2d":
            dst = view(ev["dst"])
            src = np.frombuffer(data_of(ev["data"]), dtype={1: np.uint8, 4: np.int32, 8: np.int64}[ev["dst"]["item"]])
            dst.copy_(torch.from_numpy(src.copy()).reshape(dst.shape))
        elif op == "d2d":
            view(ev["dst"]).copy_(view(ev["src"]))
        elif op == "zero":
            view(ev["dst"]).zero_()
        elif op == "d2h":
            if not check_reads:
                continue
            got = view(ev["src"]).cpu().contiguous().numpy().tobytes()
            assert _sha(got) == ev["sha"], "device -> host read %d differs from the recording" % n_d2h
            n_d2h += 1
        elif op == "call":
            name = ev["fn"]
            kinds = SPECS[name][0]
            args, outs = [], {}
            for i, (k, e) in enumerate(zip(kinds, ev["args"])):
                if k == "s":
                    args.append(e)
                elif k == "S":
                    args.append(engine.stream_ptr())
                elif k == "d":
                    args.append(C.c_void_p(addr(e["p"])))
                elif k == "T":
                    if e is None:
                        args.append(None)
                    else:
                        t = _lib.LeafTemplates.from_buffer_copy(data_of(e["tpl"]))
                        keep.append(t)
                        args.append(C.byref(t))
                elif k == "A3":
                    args.append((C.c_uint64 * 3)(*e["a3"]))
                elif isinstance(k, str) and k.startswith("hp"):
                    arr = np.array([addr(p) for p in e["ptrs"]] + [0], dtype=np.uint64)
                    keep.append(arr)
                    args.append(arr.ctypes.data_as(C.c_void_p))
                elif k[0] == "hi":
                    if e["h"] is None and e.get("null"):
                        args.append(None)
                        continue
                    buf = C.create_string_buffer(data_of(e["h"]) + b"\0" * 8)
                    keep.append(buf)
                    args.append(C.cast(buf, C.c_void_p))
                elif k[0] == "ho":
                    if e.get("null"):
                        args.append(None)
                        continue
                    buf = C.create_string_buffer(e["o"] + 8)
                    outs[i] = (buf, e["o"])
                    sig = _lib.SIGNATURES[name][1][i]
                    args.append(C.cast(buf, sig) if sig is not C.c_void_p else C.cast(buf, C.c_void_p))
            rc = getattr(lib, name)(*args)
            assert rc == ev["rc"], "%s returned %d, recorded %d: %s" % (name, rc, ev["rc"], lib.b2s_last_error())
            n_calls += 1
            for o in ev["host_out"] if check_reads else ():
                buf, nb = outs[o["arg"]]
                assert _sha(buf.raw[:nb]) == o["sha"], "call %d (%s): host output %d differs" % (n_calls, name, o["arg"])
            if check_kernels:
                for d in ev["dev_out"]:
                    loc = ev["args"][d["arg"]]["p"]
                    base = [loc[0], loc[1] + d["off"]]
                    v = {"at": base, "shape": [d["rows"], d["row_bytes"]], "strides": [d["stride"], 1], "item": 1}
                    got = view(v).cpu().contiguous().numpy().tobytes()
                    assert _sha(got) == d["sha"], "call %d (%s): device output differs from the recording" % (n_calls, name)
                    n_checked += 1
            keep.clear()
    torch.cuda.synchronize(dev) if dev.type == "cuda" else None
    return {"meta": header["meta"], "calls": n_calls, "kernel_outputs_checked": n_checked, "host_reads_checked": n_d2h,
            "allocations": len(mem)}
