"""Multi-GPU FRI prover: ONE codeword sharded over the ranks (FRI row).

code/fri.py:178-199 (prove = commit + query) with the round-0 codeword spread over G = 2^g ranks
of a torch.distributed group (NCCL on GPUs, gloo in the CPU tests).  The transcript every rank
produces is byte-identical to the one a single device (and the reference) produces.

Layout.  Split-and-fold pairs element i with i + n/2 (code/fri.py:127), so rank r starts with the
PAIR of blocks  A_r = c[r*B, (r+1)*B)  and  B_r = c[N/2 + r*B, N/2 + (r+1)*B),  B = N/(2G):

  round 0    no exchange: rank r folds [A_r | B_r] locally (the fold kernel sees a length-2B
             codeword on the coset offset*omega^(r*B); omega^(n/2) = -1 makes that exact) and
             gets block r of the next codeword plus the Merkle subtree over it;
  round k    balanced butterfly: the owners of blocks s and s + G/2 swap HALF a block each (the
             lower owner keeps the lower halves, the upper owner the upper halves), both fold
             their b/2 pairs, so every rank stays busy with 1/G of the round's work and the next
             codeword again has G blocks (of half the size, owners permuted -- _Layout.owners);
  replicate  once blocks are small (<= replicate_below elements) one all-gather gives every
             rank the whole codeword and the remaining rounds run redundantly on every rank with
             no further communication;
  trees      every block is a subtree of the round's Merkle tree; the ranks all-gather the
             subtree roots (64 B each) and each computes the few top levels itself
             (b2s_merkle_upper), so every rank knows every round's root and can run Fiat-Shamir
             (code/fri.py:120) without a broadcast;
  queries    the indices of all rounds follow from the sampled top-level indices, so all opened
             leaves and the in-subtree parts of all authentication paths are collected with ONE
             all-reduce (owners contribute, everybody else zeros); the top levels come from the
             replicated top trees.

All ranks execute the same sequence of collectives (SPMD) and end with identical proof streams.
"""
import numpy as np
import torch
import torch.distributed as dist

from .glue import DeviceCodeword, NodeView, prefetch_openings

P = 18446744069414584321


def scatter_pair_blocks(planes, rank, world):
    """host helper: the (A_r, B_r) blocks rank `rank` owns of a full codeword given as (3, N) planes"""
    a = np.asarray(planes, dtype=np.uint64)
    n = a.shape[1]
    blk = n // (2 * world)
    lo = rank * blk
    return (np.ascontiguousarray(a[:, lo:lo + blk]), np.ascontiguousarray(a[:, n // 2 + lo:n // 2 + lo + blk]))


def residues_to_pair_blocks(planes, group=None):
    """The residue-class shard of a codeword that dist.shard_coset_evaluate produces (rank r: c[G t + r], (q, N/G))
    -> the pair of blocks DistFri.prove starts from (A_r, B_r, each (q, N/2G)): one all-to-all of the codeword
    ("the LDE row's output is residue-sharded; the FRI row wants index ranges").  Needs N >= 2 G^2."""
    world = dist.get_world_size(group)
    q, n_loc = planes.shape
    if world == 1:
        return planes[:, :n_loc // 2].contiguous(), planes[:, n_loc // 2:].contiguous()
    blk = n_loc // 2            # = N / 2G, the block length
    assert blk % world == 0, "codeword too short for this many ranks"
    u = blk // world
    # t = h * (N/2G) + d * (B/G) + u'  <->  index h * N/2 + d * B + (G u' + r): block d of half h, offset G u' + r
    send = planes.reshape(q, 2, world, u).permute(2, 0, 1, 3).contiguous()  # [destination d][plane][half][u']
    recv = torch.empty_like(send)                                           # [source r][plane][half][u']
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    blocks = recv.permute(1, 2, 3, 0).reshape(q, 2, blk)                    # offset G u' + r
    return blocks[:, 0].contiguous(), blocks[:, 1].contiguous()


class _Layout:
    """where the leaves of one round live: subtree s covers leaves [s*blk, (s+1)*blk) and is slot
    `slot` of rank `rank`, (rank, slot) = owner(s)"""

    def __init__(self, n, blk, world, owners=None):
        self.n, self.blk, self.world = n, blk, world
        self.subtrees = n // blk
        self.paired = owners is None  # round 0: subtrees 0..G-1 are the A blocks, G..2G-1 the B blocks
        self.owners = owners          # later rounds: one block per rank, owners[s] = rank of subtree s

    def owner(self, s):
        if self.paired:
            return s % self.world, s // self.world
        return self.owners[s], 0


class DistCodeword(DeviceCodeword):
    """glue.DeviceCodeword over a codeword whose blocks live on several ranks; prefetch() is a
    collective (every rank must call it with the same indices, which SPMD Fiat-Shamir ensures)"""

    def __init__(self, df, layout, local_planes, xfield):
        self._glue = df.glue
        self._df = df
        self._layout = layout
        self._local = local_planes  # slot -> (3, blk) device tensor (None on ranks that own nothing)
        self._field = xfield
        self._n = layout.n
        self._cache = {}

    # prefetch = wanted() -> local() -> [sum over ranks] -> fill(); DistFri.prefetch_queries runs the
    # middle step ONCE for all trees and codewords of the query phase
    def local(self, need):
        """(len(need), 3) uint64: the elements this rank owns, zeros elsewhere"""
        lay, df = self._layout, self._df
        vals = np.zeros((len(need), 3), dtype=np.uint64)
        mine = {}
        for pos, i in enumerate(need):
            r, slot = lay.owner(i // lay.blk)
            if r == df.rank:
                mine.setdefault(slot, []).append((pos, i % lay.blk))
        for slot, items in mine.items():
            vals[[pos for pos, _ in items]] = df.eng.gather(self._local[slot], [j for _, j in items])
        return vals

    def prefetch(self, indices):
        need = self.wanted(indices)  # wanted() / fill() are the base class's
        if need:
            self.fill(need, self._df.sum_over_ranks(self.local(need)))

    def materialize(self):
        self.prefetch(range(self._n))
        return [self._cache[i] for i in range(self._n)]


class DistNodeView(NodeView):
    """glue.NodeView over a tree whose bottom levels are per-rank subtrees and whose top levels
    (above the subtree roots) are replicated on every rank"""

    def __init__(self, df, layout, local_nodes, top):
        self._df = df
        self._layout = layout
        self._local = local_nodes  # slot -> (2 blk, 64) uint8 heap of the local subtree
        self._npo2 = self._n = layout.n
        self._cache = dict(top)    # heap index -> bytes for every node of index < 2 * subtrees

    def _fetch_all(self):
        raise NotImplementedError("a sharded tree is only opened, never listed")

    def __getitem__(self, k):
        if isinstance(k, slice):
            self._fetch_all()
        if k < 0:
            k += len(self)
        if not 0 <= k < len(self):
            raise IndexError("list index out of range")
        v = self._cache.get(k)
        if v is None:
            if k == 0:
                from hashlib import blake2b
                v = self._cache[0] = blake2b(self._ZERO32 + self[1]).digest()
            else:
                raise KeyError("node %d of a sharded tree was not prefetched (prefetch_paths is collective)" % k)
        return v

    def wanted(self, indices):
        lay = self._layout
        low = lay.blk.bit_length() - 1  # levels inside a subtree
        n = lay.n
        return [i for i in dict.fromkeys(indices) if 0 <= i < n
                and any(((n | i) >> j) ^ 1 not in self._cache for j in range(low))]

    def local(self, need):
        """(len(need), low, 64) uint8: the in-subtree part of the paths this rank owns, zeros elsewhere"""
        lay, df = self._layout, self._df
        low = lay.blk.bit_length() - 1
        buf = np.zeros((len(need), low, 64), dtype=np.uint8)
        mine = {}
        for pos, i in enumerate(need):
            r, slot = lay.owner(i // lay.blk)
            if r == df.rank:
                mine.setdefault(slot, []).append((pos, i % lay.blk))
        for slot, items in mine.items():
            paths = df.eng.merkle_open(self._local[slot], [j for _, j in items])
            for (pos, _), path in zip(items, paths):
                buf[pos] = np.frombuffer(b"".join(path), dtype=np.uint8).reshape(low, 64)
        return buf

    def fill(self, need, buf):
        n = self._layout.n
        for pos, i in enumerate(need):
            k = n | i
            for j in range(buf.shape[1]):
                sib = (k >> j) ^ 1
                if sib not in self._cache:
                    self._cache[sib] = buf[pos, j].tobytes()

    def prefetch_paths(self, indices, depth):
        need = self.wanted(indices)
        if need:
            self.fill(need, self._df.sum_over_ranks(self.local(need)))


class DistFri:
    def __init__(self, glue, group=None):
        self.glue = glue
        self.eng = glue.engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert self.world & (self.world - 1) == 0, "power-of-two number of ranks"
        self.exchanged_bytes = 0  # payload this rank sent or received point to point (for reports)

    # ---- collectives -----------------------------------------------------------------------
    def sum_over_ranks(self, arr):
        """numpy array, non-zero on exactly one rank per entry -> the same array on every rank"""
        if self.world == 1:
            return arr
        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(self.eng.device)  # bytes: no carries to worry about
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().view(a.dtype).reshape(a.shape)

    def prefetch_queries(self, requests):
        """everything the query phase will open -- requests = [(DistCodeword | DistNodeView, indices)] --
        with ONE all-reduce instead of one per tree and round"""
        jobs = [(obj, obj.wanted(idx)) for obj, idx in requests]
        jobs = [(obj, need) for obj, need in jobs if need]
        if not jobs:
            return
        parts = [np.ascontiguousarray(obj.local(need)) for obj, need in jobs]
        flat = np.concatenate([p.view(np.uint8).reshape(-1) for p in parts])
        flat = self.sum_over_ranks(flat)
        pos = 0
        for (obj, need), p in zip(jobs, parts):
            obj.fill(need, flat[pos:pos + p.nbytes].view(p.dtype).reshape(p.shape))
            pos += p.nbytes

    def _gather_roots(self, mine, per_rank):
        """all-gather of `per_rank` 64-byte digests per rank -> (world, per_rank, 64) uint8 on the device"""
        out = torch.empty(self.world * per_rank * 64, dtype=torch.uint8, device=self.eng.device)
        dist.all_gather_into_tensor(out, mine.contiguous().view(-1), group=self.group)
        return out.view(self.world, per_rank, 64)

    def _tree(self, Merkle, layout, local_planes, local_nodes, xfield):
        """Merkle object of one round: subtree roots exchanged, top levels computed on every rank"""
        eng = self.eng
        slots = 2 if layout.paired else 1
        mine = torch.zeros((slots, 64), dtype=torch.uint8, device=eng.device)
        for s in range(slots):
            if local_nodes[s] is not None:
                mine[s] = local_nodes[s][1]
        roots = self._gather_roots(mine, slots)  # [rank][slot]
        S = layout.subtrees
        heap = torch.zeros((2 * S, 64), dtype=torch.uint8, device=eng.device)
        if layout.paired:
            heap[S:] = roots.transpose(0, 1).reshape(S, 64)  # subtree s = slot s // G of rank s % G
        else:
            heap[S:] = roots[layout.owners, 0]
        if S > 1:
            eng.merkle_upper(heap)
        raw = eng.download_bytes(heap)
        top = {k: raw[64 * k:64 * k + 64] for k in range(1, 2 * S)}
        tree = Merkle.__new__(Merkle)
        tree.num_leafs = layout.n
        tree.depth = layout.n.bit_length() - 1
        tree.leafs = DistCodeword(self, layout, local_planes, xfield)
        tree.nodes = DistNodeView(self, layout, local_nodes, top)
        return tree

    def _butterfly(self, own, layout):
        """round k >= 1: swap half a block with the owner of the paired subtree; returns the
        [low | high] pair buffer this rank folds, the index of the subtree it produces and the
        owners of the next round's subtrees"""
        S, b = layout.subtrees, layout.blk
        s = layout.owners.index(self.rank)
        h = b // 2
        lower = s < S // 2
        partner = layout.owners[s + S // 2] if lower else layout.owners[s - S // 2]
        send = own[:, h:].contiguous() if lower else own[:, :h].contiguous()
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, group=self.group, group_peer=partner),
               dist.P2POp(dist.irecv, recv, group=self.group, group_peer=partner)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        self.exchanged_bytes += send.numel() * 8
        if lower:
            pair, t = torch.cat([own[:, :h], recv], dim=1), 2 * s
        else:
            pair, t = torch.cat([recv, own[:, h:]], dim=1), 2 * (s - S // 2) + 1
        owners = [layout.owners[(u >> 1) + (u & 1) * (S // 2)] for u in range(S)]
        return pair, t, owners

    def _replicate(self, own, layout):
        """all-gather of the blocks: the whole codeword of this round on every rank"""
        b = layout.blk
        out = torch.empty(self.world * 3 * b, dtype=own.dtype, device=own.device)
        dist.all_gather_into_tensor(out, own.contiguous().view(-1), group=self.group)
        return out.view(self.world, 3, b)[layout.owners].permute(1, 0, 2).reshape(3, layout.n).contiguous()

    # ---- code/fri.py:91-139 + :178-199 -------------------------------------------------------
    def prove(self, fri, block_a, block_b, proof_stream, Merkle, replicate_below=1 << 12):
        """block_a / block_b: this rank's (3, B) device planes of the round-0 codeword (see module
        docstring; scatter_pair_blocks() cuts them from a full codeword).  Pushes exactly what
        Fri.prove pushes and returns its top_level_indices -- on every rank."""
        glue, eng, G, rank = self.glue, self.eng, self.world, self.rank
        xfield = fri.field
        N = fri.domain.length
        blk = N // (2 * G)
        assert blk >= 1 and blk * 2 * G == N and block_a.shape == (3, blk) and block_b.shape == (3, blk), \
            "initial codeword length does not match length of initial codeword"
        num_rounds = fri.num_rounds()
        omega = glue.base_value(fri.domain.omega)
        offset = glue.base_value(fri.domain.offset)
        tpl = glue.xfe_templates(xfield)

        layout = _Layout(N, blk, G)  # None once the codeword is replicated
        planes = [block_a, block_b]
        nodes = [eng.merkle_field(block_a, tpl), eng.merkle_field(block_b, tpl)]
        n = N
        trees, codewords = [], []
        for r in range(num_rounds):
            assert pow(omega, n - 1, P) == pow(omega, P - 2, P), "error in commit: omega does not have the right order!"
            if layout is not None:
                tree = self._tree(Merkle, layout, planes, nodes, xfield)
            else:
                tree = Merkle.__new__(Merkle)
                cache = DeviceCodeword(glue, planes[0], xfield)
                glue.merkle_build(tree, cache, device_planes=planes[0], device_nodes=nodes[0], leaf_cache=cache)
            root = tree.root()
            if r > 0:
                proof_stream.push(root)
            if r == num_rounds - 1:
                break
            alpha = xfield.sample(proof_stream.prover_fiat_shamir())
            codewords.append(tree.leafs)
            trees.append(tree)
            a = [c.value for c in alpha.polynomial.coefficients]
            a += [0] * (3 - len(a))
            # the [low | high] pair buffer this rank folds and where its first pair sits in the round
            if layout is None:
                pair, first, layout_next = planes[0], 0, None
            elif layout.paired:
                pair, first = torch.cat([planes[0], planes[1]], dim=1), rank * blk
                layout_next = _Layout(n // 2, blk, G, owners=list(range(G)))
            elif G > 1 and layout.blk >= 2 and layout.blk > replicate_below:
                pair, t, owners = self._butterfly(planes[0], layout)
                first = t * (layout.blk // 2)
                layout_next = _Layout(n // 2, layout.blk // 2, G, owners=owners)
            else:
                pair, first, layout_next = self._replicate(planes[0], layout), 0, None
            nxt, nn = eng.fri_fold(pair, a, offset * pow(omega, first, P) % P, omega, tpl)
            planes, nodes, layout = [nxt], [nn], layout_next
            n //= 2
            omega = omega * omega % P
            offset = offset * offset % P
        last = tree.leafs.materialize()
        proof_stream.push(last)
        codewords.append(last)

        # code/fri.py:186-199
        top_level_indices = fri.sample_indices(proof_stream.prover_fiat_shamir(), len(codewords[1]),
                                               len(codewords[-1]), fri.num_colinearity_tests)
        # the indices of every round follow from the top-level ones: fetch all remote openings at once
        sq = fri.num_colinearity_tests
        wanted, idx = {}, [i for i in top_level_indices]
        for i in range(len(trees)):
            half = len(codewords[i]) // 2
            idx = [index % half for index in idx]
            wanted.setdefault(i, []).extend(idx[:sq] + [j + half for j in idx[:sq]])
            if i + 1 < len(trees):
                wanted.setdefault(i + 1, []).extend(idx[:sq])
        sharded = [i for i in wanted if isinstance(trees[i].nodes, DistNodeView)]
        self.prefetch_queries([(trees[i].leafs, wanted[i]) for i in sharded] +
                              [(trees[i].nodes, wanted[i]) for i in sharded])
        # the replicated tail rounds are plain device trees: one local call for all of them
        prefetch_openings(eng, trees, [wanted.get(i, []) for i in range(len(trees))])
        indices = [i for i in top_level_indices]
        for i in range(len(trees) - 1):
            indices = [index % (len(codewords[i]) // 2) for index in indices]
            glue.fri_query(fri, trees[i], trees[i + 1], indices, proof_stream)
        indices = [index % len(codewords[-1]) for index in indices]
        glue.fri_query_last(fri, trees[-1], codewords[-1], indices, proof_stream)
        return top_level_indices
