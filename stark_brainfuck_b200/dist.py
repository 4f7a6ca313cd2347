"""Multi-GPU: one process per GPU (torch.distributed: NCCL over NVLink 5 / NVSwitch; gloo in CPU
tests).  What shards, and how:

  * batches of columns / planes / whole proofs are independent units: shard_units() splits them
    over ranks, no data-path collective (this is what bench.py --gpus N measures, weak scaling);
  * ONE transform of length n = n1*n2 shards with ONE exchange (four-step): DistNTT below.
      input  (rank g): planes P[j1_local][j2] = x[(g*q + j1_local) + n1*j2],  q = n1/G
      step 1 local length-n2 transforms (root omega^n1)            -- b2s_ntt, batched planes
      step 2 twiddle omega^(j1*k2) + transpose + placement         -- b2s_dist_twiddle_transpose
      step 3 exchange: NCCL all-to-all, or none at all when step 2 stored straight into the
             peers' buffers over NVLink ("p2p": torch symmetric memory supplies the pointers)
      step 4 local length-n1 transforms (root omega^n2)            -- b2s_ntt, batched planes
      output (rank g): planes D[k2_local][k1] = X[k1*n2 + g*(n2/G) + k2_local]
    The output has the input's layout with n1 and n2 swapped, so the inverse is the same
    routine (inverse=True, log_n1 = log2(n2)).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

P = 18446744069414584321


def shard_units(n_units, rank, world):
    """contiguous range of independent units (planes, columns, proofs) owned by `rank`"""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def scatter_columns(x, log_n, log_n1, rank, world):
    """host helper: the planes rank `rank` owns of a full natural-order vector x (numpy uint64)"""
    n1, n2 = 1 << log_n1, 1 << (log_n - log_n1)
    q = n1 // world
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64).reshape(n2, n1).T[rank * q:(rank + 1) * q])


def assemble_output(parts, log_n, log_n1):
    """host helper: natural-order result from the per-rank outputs D_g (cpp, n1) in rank order"""
    n1, n2 = 1 << log_n1, 1 << (log_n - log_n1)
    d = np.concatenate(parts, axis=0)  # (n2, n1): d[k2][k1] = X[k1*n2 + k2]
    return np.ascontiguousarray(d.T).reshape(n1 * n2)


def shard_coset_evaluate(engine, coeffs, log_n, omega, offset, rank, world, group=None, gather_coefficients=False):
    """Coset evaluation / LDE of ONE polynomial over the ranks (code/fri.py:26-37,
    code/ntt.py:164-168): rank r computes the residue class r of the output,
        out[G t + r] = sum_j (c_j (offset omega^r)^j) (omega^G)^(j t),      t < n / G,
    a size-n/G coset transform of the coefficients with offset offset*omega^r and root omega^G -- no exchange.
    With up to 2 n/G coefficients the rank transforms on the twice finer coset and keeps the even outputs; with
    more they are scaled by (offset omega^r)^j and folded modulo n/G first (one polynomial per call then).
    coeffs: (q, m) planes on the device (a base-field polynomial per plane; an extension-field one is three), replicated on every
    rank -- or, with gather_coefficients=True, this rank's contiguous 1/G of them (one all-gather of the small
    coefficient vector is then the only collective).  Returns (q, n/G) planes: element t is output G t + r."""
    G = world
    assert G & (G - 1) == 0, "power-of-two number of ranks"
    if gather_coefficients and G > 1:
        q, m_loc = coeffs.shape
        full = torch.empty(G * q * m_loc, dtype=coeffs.dtype, device=coeffs.device)
        dist.all_gather_into_tensor(full, coeffs.contiguous().view(-1), group=group)
        coeffs = full.view(G, q, m_loc).permute(1, 0, 2).reshape(q, G * m_loc).contiguous()
    q, m = coeffs.shape
    log_g = G.bit_length() - 1
    n_loc = 1 << (log_n - log_g)
    w_g = pow(omega, G, P)
    off_r = offset * pow(omega, rank, P) % P
    if m <= n_loc:
        return engine.ntt(coeffs, log_n - log_g, w_g, offset=off_r)  # any number of planes: one batched call
    if m <= 2 * n_loc:
        # Twice as many coefficients as local points (expansion factor G/2): the transform of length 2 n/G on the
        # coset offset*omega^r with root omega^(G/2) has the wanted values at its even outputs.  It computes twice
        # the outputs the rank keeps, but takes every plane of a trace in ONE call, where scaling and folding
        # each polynomial separately (below) is latency-bound (46 planes on 8 GPUs: 2.7 ms against 0.9 on one).
        full = engine.ntt(coeffs, log_n - log_g + 1, pow(omega, G // 2, P), offset=off_r)
        return full[:, ::2].contiguous()
    assert q in (1, 3)
    # more coefficients than local points: x^(n/G) = (offset omega^r)^(n/G) on this residue class after scaling,
    # i.e. scale first, then add the chunks of n/G coefficients on top of each other (b2s_combination with unit
    # weights does the modular sum), then a plain transform
    chunks = -(-m // n_loc)
    scaled = engine.scale(coeffs, off_r if q == 1 else [off_r, 0, 0])
    if chunks * n_loc != m:
        padded = torch.zeros((q, chunks * n_loc), dtype=scaled.dtype, device=scaled.device)
        padded[:, :m] = scaled
        scaled = padded
    cols = [scaled[:, k * n_loc:(k + 1) * n_loc] for k in range(chunks)]
    ones = np.zeros((chunks, 3), dtype=np.uint64)
    ones[:, 0] = 1
    folded = engine.combination(cols, ones, np.zeros((chunks, 3), dtype=np.uint64), [0] * chunks, n_loc, 1, 1)
    return engine.ntt(folded[:q], log_n - log_g, w_g)


def columns_to_rows(planes, group=None):
    """Column-sharded codewords -> row-sharded ones: every rank holds the SAME number of whole
    planes (cols_local, N) -- what plane-parallel transforms produce -- and ends up with the index range
    [r N/G, (r+1) N/G) of ALL planes, (G * cols_local, N/G), plane order = rank order: one all-to-all of
    8 N cols_local (G-1)/G bytes per rank.  This is the layout zipped-row Merkle leaves and the sharded FRI
    prover (dist_fri) want."""
    world = dist.get_world_size(group)
    c, n = planes.shape
    assert n % world == 0
    if world == 1:
        return planes
    blk = n // world
    send = planes.reshape(c, world, blk).permute(1, 0, 2).contiguous()  # [dest rank][plane][index]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    return recv.view(world * c, blk)                                    # [source rank][plane] -> plane-major


def shard_eval_points(engine, coeffs, points, rank, world):
    """code/univariate.py:145-154 on arbitrary points, point-range sharded: every rank
    evaluates its contiguous share of the points, no collective.  Returns (planes, share of the points)."""
    share = shard_units(points.shape[1], rank, world)
    return engine.eval_points(coeffs, points[:, share.start:share.stop].contiguous())


def assemble_residues(parts):
    """host helper: natural-order output from the per-rank residue classes (q, n/G), rank order"""
    G = len(parts)
    q, n_loc = parts[0].shape
    out = np.empty((q, G * n_loc), dtype=parts[0].dtype)
    for r, p_ in enumerate(parts):
        out[:, r::G] = p_
    return out


class DistNTT:
    def __init__(self, engine, group=None, exchange="nccl"):
        self.eng = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert self.world & (self.world - 1) == 0, "power-of-two number of ranks"
        self.exchange = exchange
        self._symm = {}  # (elements) -> (tensor, handle)

    def _symm_buffer(self, numel):
        import torch.distributed._symmetric_memory as symm_mem
        ent = self._symm.get(numel)
        if ent is None:
            t = symm_mem.empty((numel,), dtype=torch.int64, device=self.eng.device)
            h = symm_mem.rendezvous(t, group=self.group if self.group is not None else dist.group.WORLD)
            ent = (t, h)
            self._symm[numel] = ent
        return ent

    def transform(self, planes, log_n, omega, inverse=False, log_n1=None):
        """planes: (q, n2) int64 tensor on the engine's device.  Returns (n2/G, n1)."""
        eng, G, g = self.eng, self.world, self.rank
        if log_n1 is None:
            log_n1 = log_n // 2
        log_n2 = log_n - log_n1
        n1, n2 = 1 << log_n1, 1 << log_n2
        q, cpp = n1 // G, n2 // G
        assert q >= 1 and cpp >= 1, "more ranks than rows/columns"
        assert tuple(planes.shape) == (q, n2)
        n = 1 << log_n
        assert pow(omega, n, P) == 1 and (n < 2 or pow(omega, n // 2, P) != 1), \
            "primitive root must be nth root of unity"
        w1 = pow(omega, n1, P)  # order n2
        w2 = pow(omega, n2, P)  # order n1
        wt = pow(omega, P - 2, P) if inverse else omega
        a = eng.ntt(planes, log_n2, w1, inverse=inverse) if log_n2 else planes
        lib = eng.lib
        if self.exchange == "p2p" and G > 1:
            buf, hdl = self._symm_buffer(cpp * n1)
            hdl.barrier()  # every rank is done reading the previous contents
            ptrs = (C.c_void_p * G)(*[int(hdl.buffer_ptrs[r]) for r in range(G)])
            eng.check(lib.b2s_dist_twiddle_transpose(C.c_void_p(a.data_ptr()), a.stride(0), q, n2, g * q, wt, 1, ptrs,
                                                     G, n1, g * q, eng.stream_ptr()))
            hdl.barrier()  # all peers' stores have landed
            c = buf.view(cpp, n1)
        else:
            send = torch.empty(G * cpp * q, dtype=torch.int64, device=eng.device)
            ptrs = (C.c_void_p * G)(*[send.data_ptr() + 8 * r * cpp * q for r in range(G)])
            eng.check(lib.b2s_dist_twiddle_transpose(C.c_void_p(a.data_ptr()), a.stride(0), q, n2, g * q, wt, 1, ptrs,
                                                     G, q, 0, eng.stream_ptr()))
            if G > 1:
                recv = torch.empty_like(send)
                dist.all_to_all_single(recv, send, group=self.group)
                c = torch.empty((cpp, n1), dtype=torch.int64, device=eng.device)
                eng.check(lib.b2s_block_permute(C.c_void_p(recv.data_ptr()), C.c_void_p(c.data_ptr()), G, cpp, q,
                                                eng.stream_ptr()))
            else:
                c = send.view(cpp, n1)
        return eng.ntt(c, log_n1, w2, inverse=inverse) if log_n1 else c
