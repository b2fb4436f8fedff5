"""CPU stand-in for one rank of the multi-GPU BPE trainer (TEST INFRASTRUCTURE).

It speaks the same three-phase protocol as subword_tokenizers_b200.device.CudaTrainEngine
(select | merge | update, exchange tensors init_counts / cand / cand_gather / delta) so that the PRODUCT's
loop `device.run_training_loop` -- including its torch.distributed collectives and `shard_types` -- can be
exercised with world_size 2 over gloo on a box without GPUs.  The arithmetic mirrors the kernels of
csrc/bpe_train.cu in plain numpy/Python; it is never used by the product."""
import numpy as np
import torch

NOPOS = np.uint64(0xFFFFFFFFFFFFFFFF)


class NumpyTrainEngine:
    def __init__(self, syms, off, freq, n_alpha, max_vocab, initial_vocab, slot_base, rank, world_size, record_cap=64):
        self.words = [list(map(int, syms[int(off[i]):int(off[i + 1])])) for i in range(len(off) - 1)]
        self.starts = [int(o) for o in off[:-1]]
        self.freq = [int(f) for f in freq]
        self.n_alpha, self.max_vocab, self.slot_base = n_alpha, max_vocab, slot_base
        self.rank, self.world, self.record_cap = rank, world_size, record_cap
        self.vmax = max(max_vocab, n_alpha) + 2
        self.init_counts = torch.zeros(n_alpha * n_alpha, dtype=torch.int64)
        self.cand = torch.zeros(2, dtype=torch.int64)
        self.cand_gather = torch.zeros(2 * world_size, dtype=torch.int64)
        self.delta = torch.zeros(2 * self.vmax + 2, dtype=torch.int64)
        self.table = {}
        self.strs = [(c,) for c in range(n_alpha)]
        self.str_ids = {}
        self.vocab_size = initial_vocab
        self.halt = 0
        self.records = []
        self.n_total = 0
        self.cur = None
        self.n_tied = 0
        self.max_count = 0
        self.cand_key = 0

    def count_local(self):
        c = self.init_counts.numpy()
        for w, f in zip(self.words, self.freq):
            for a, b in zip(w, w[1:]):
                c[a * self.n_alpha + b] += f

    def build_table(self):
        c = self.init_counts.numpy()
        for i in np.nonzero(c)[0]:
            self.table[(int(i) // self.n_alpha, int(i) % self.n_alpha)] = int(c[i])

    def steps(self, n):
        for _ in range(n):
            self.select(); self.cand_gather[:2] = self.cand; self.merge(); self.update()

    def select(self):
        self.cur = None
        cand = self.cand.numpy().view(np.uint64)
        cand[0], cand[1] = NOPOS, NOPOS
        if self.halt:
            return
        best = max((c for c in self.table.values() if c > 0), default=0)
        tied = sorted(k for k, c in self.table.items() if c == best and c > 0)
        self.max_count, self.n_tied = best, len(tied)
        if self.vocab_size >= self.max_vocab:
            self.halt = 1
        elif best <= 0:
            self.halt = 2
        elif len(self.records) >= self.record_cap:
            self.halt = 4
        if self.halt:
            return
        self.cand_key = (tied[0][0] << 32) | tied[0][1]
        if self.n_tied <= 1:
            cand[1] = self.cand_key
            return
        tied = set(tied)
        for w, s0 in zip(self.words, self.starts):          # ascending slot order, first hit wins
            for i, pair in enumerate(zip(w, w[1:])):
                if pair in tied:
                    cand[0] = np.uint64(self.slot_base + s0 + i)
                    cand[1] = np.uint64((pair[0] << 32) | pair[1])
                    return

    def merge(self):
        if self.halt:
            return
        g = self.cand_gather.numpy().view(np.uint64)
        if self.n_tied <= 1:
            key = self.cand_key
        else:
            r = int(np.argmin(g[0::2]))
            assert g[2 * r] != NOPOS
            key = int(g[2 * r + 1])
        a, b = key >> 32, key & 0xFFFFFFFF
        s = self.strs[a] + self.strs[b]
        z = self.str_ids.get(s)
        if z is None:
            z = len(self.strs); self.strs.append(s); self.str_ids[s] = z; self.vocab_size += 1
        self.records.append((a, b, z, self.max_count)); self.n_total += 1
        self.cur = (a, b, z)
        d = self.delta.numpy()
        L, R = d[:self.vmax], d[self.vmax:2 * self.vmax]
        for wi, w in enumerate(self.words):
            f, out, i, last_merge = self.freq[wi], [], 0, False
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    if out:
                        if last_merge: d[2 * self.vmax] += f
                        else: L[out[-1]] += f
                    d[2 * self.vmax + 1] += f
                    out.append(z); last_merge = True; i += 2
                else:
                    if out and last_merge: R[w[i]] += f
                    out.append(w[i]); last_merge = False; i += 1
            self.words[wi] = out           # slot positions keep their order; dead slots do not matter here

    def update(self):
        if self.halt or self.cur is None:
            return
        a, b, z = self.cur
        d = self.delta.numpy()
        L, R = d[:self.vmax], d[self.vmax:2 * self.vmax]
        t = self.table
        for x in np.nonzero(L)[0]:
            x = int(x); t[(x, a)] = t.get((x, a), 0) - int(L[x]); t[(x, z)] = t.get((x, z), 0) + int(L[x])
        for y in np.nonzero(R)[0]:
            y = int(y); t[(b, y)] = t.get((b, y), 0) - int(R[y]); t[(z, y)] = t.get((z, y), 0) + int(R[y])
        zz, m = int(d[2 * self.vmax]), int(d[2 * self.vmax + 1])
        if zz:
            t[(b, a)] = t.get((b, a), 0) - zz; t[(z, z)] = t.get((z, z), 0) + zz
        t[(a, b)] = t.get((a, b), 0) - m
        d[:] = 0

    def read(self):
        rec = self.records; self.records = []
        state = {"halt": self.halt, "n_recorded": len(rec), "n_merges_total": self.n_total, "vocab_size": self.vocab_size,
                 "table_cap": 0}
        if self.halt == 4:
            self.halt = 0
        arr = lambda k, dt: np.array([r[k] for r in rec], dtype=dt)
        return state, arr(0, np.uint32), arr(1, np.uint32), arr(2, np.uint32), arr(3, np.int64)

    def grow_table(self, cap):
        raise AssertionError("the numpy engine never asks to grow")
