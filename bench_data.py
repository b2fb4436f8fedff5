"""Deterministic synthetic inputs of the benchmarks and the large parity tests.

* ``train5k_types``      the 22,971 word types of data/train-5K.json (BERT pre-tokenized, lower-cased), first-occurrence order
* ``synth_type_table``   N synthetic word types: concatenations of 1-4 units drawn from the merged strings of the reference's
                         pretrained BPE model plus single letters, clipped to the measured length histogram of train-5K
                         (mean 8.2, max 22), duplicates rejected -- generated with numpy only (10 M types in ~20 s)
* ``ZipfStream``         i.i.d. word stream over a type table with integer weights max(1, floor(C / rank)); the draw sequence
                         comes from PCG64 in fixed chunks, so any prefix is reproducible on the CPU and on the GPU

Nothing here is on the product path; it only manufactures inputs.
"""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")
DRAW_CHUNK = 1 << 24


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt", encoding="utf-8") as f:
        return json.load(f)


def train5k_types():
    from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
    pre = make_hf_tokenizer().backend_tokenizer.pre_tokenizer
    corpus = load_golden("train-5K.json.gz")
    return list(dict.fromkeys(w for s in corpus for w, _ in pre.pre_tokenize_str(s.lower())))


def _units():
    merges = load_golden("pretrained_bpe_merges.json.gz")
    units = sorted({a + b for a, b in merges})
    alphabet = sorted({c for u in units for c in u})
    units = units + alphabet
    lens = np.fromiter((len(u) for u in units), dtype=np.int64, count=len(units))
    off = np.zeros(len(units) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    arena = np.frombuffer("".join(units).encode("utf-32-le"), dtype=np.uint32).astype(np.uint16)
    return arena, off, lens


def synth_type_table(n_types: int, seed: int, max_len: int = 22):
    """-> (mat uint16 [n_types, max_len] code points padded with 0, lens int64 [n_types]); rows are distinct."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u_arena, u_off, u_len = _units()
    nu = len(u_len)
    pos = np.arange(max_len, dtype=np.int64)
    pw = np.uint64(0x9E3779B97F4A7C15) ** np.arange(1, max_len + 1, dtype=np.uint64)      # wraps mod 2^64
    rows, seen, have = [], np.zeros(0, dtype=np.uint64), 0
    u_len32, u_off32, pos32 = u_len.astype(np.int32), u_off.astype(np.int32), pos.astype(np.int32)
    while have < n_types:
        m = int(min(2_000_000, max(1 << 16, (n_types - have) * 5 // 4)))
        k = rng.integers(1, 5, size=m)
        idx = rng.integers(0, nu, size=(m, 4)).astype(np.int32)
        tgt = np.clip(np.rint(rng.normal(8.2, 3.0, size=m)), 2, max_len).astype(np.int32)     # train-5K: mode 7-8, mean 8.18, max 22
        ul = u_len32[idx] * (np.arange(4, dtype=np.int32)[None, :] < k[:, None])
        cum = np.cumsum(ul, axis=1, dtype=np.int32)                                          # [m, 4] inclusive
        total = np.minimum(cum[:, 3], tgt)
        # unit that holds position j, the unit's first position, and the source index of the character
        src = np.empty((m, max_len), dtype=np.int32)
        c0, c1, c2 = cum[:, 0:1], cum[:, 1:2], cum[:, 2:3]
        p = pos32[None, :]
        in0, in1, in2 = p < c0, p < c1, p < c2
        uo = u_off32[idx]                                                                    # [m, 4] first character of every unit
        src[:] = np.where(in0, uo[:, 0:1] + p, np.where(in1, uo[:, 1:2] + (p - c0), np.where(in2, uo[:, 2:3] + (p - c1), uo[:, 3:4] + (p - c2))))
        valid = p < total[:, None]
        np.putmask(src, ~valid, 0)
        mat = u_arena[src]
        np.putmask(mat, ~valid, 0)
        h = (mat.astype(np.uint64) * pw[None, :]).sum(axis=1, dtype=np.uint64) + total.astype(np.uint64)
        _, first = np.unique(h, return_index=True)
        keep = np.zeros(m, dtype=bool)
        keep[first] = True
        if len(seen):
            keep &= ~np.isin(h, seen)
        sel = np.flatnonzero(keep)[: n_types - have]
        rows.append((mat[sel], total[sel].astype(np.int64)))
        seen = np.concatenate([seen, h[sel]])
        have += len(sel)
    mat = np.concatenate([r[0] for r in rows])
    lens = np.concatenate([r[1] for r in rows])
    return mat, lens


def table_to_cps(mat, lens):
    """-> (cps uint32 concatenated, off uint64 [n + 1])."""
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens.astype(np.uint64), out=off[1:])
    mask = np.arange(mat.shape[1])[None, :] < lens[:, None]
    return mat[mask].astype(np.uint32), off


def table_to_utf8(mat, lens):
    """-> (arena uint8, off int64 [n + 1]) UTF-8 bytes of every type (code points < 0x800)."""
    mask = np.arange(mat.shape[1])[None, :] < lens[:, None]
    two = (mat >= 0x80) & mask
    nb = (mask.astype(np.int64) + two).sum(axis=1)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(nb, out=off[1:])
    cps = mat[mask].astype(np.uint32)
    big = cps >= 0x80
    out = np.empty(int(off[-1]), dtype=np.uint8)
    p = np.cumsum(1 + big) - (1 + big)                      # byte position of every character
    out[p] = np.where(big, 0xC0 | (cps >> 6), cps).astype(np.uint8)
    out[p[big] + 1] = (0x80 | (cps[big] & 0x3F)).astype(np.uint8)
    return out, off


def table_to_strings(mat, lens):
    return ["".join(map(chr, mat[i, :lens[i]])) for i in range(len(lens))]


def zipf_freqs(n_types: int, seed: int, c_factor: int = 20):
    """Training frequencies: rank r has frequency max(1, floor(C / r)), C = c_factor * n_types; ranks are assigned to the types by
    a seeded permutation (the type ORDER, which drives the trainer's tie-break, is the table order)."""
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    order = rng.permutation(n_types)
    freq = np.zeros(n_types, dtype=np.int64)
    freq[order] = np.maximum(1, (c_factor * n_types) // np.arange(1, n_types + 1))
    return freq


class ZipfStream:
    """i.i.d. word stream over a type table (UTF-8 arena + offsets), rank r drawn with integer weight max(1, floor(C / r))."""

    def __init__(self, t_arena, t_off, zipf_c: int, seed: int, permute: bool = True):
        n = len(t_off) - 1
        rng = np.random.Generator(np.random.PCG64(seed))
        order = rng.permutation(n) if permute else np.arange(n)
        self.order = order                                   # rank r is type order[r]
        weights = np.maximum(1, zipf_c // np.arange(1, n + 1)).astype(np.int64)
        self.lut = np.repeat(order.astype(np.int32), weights)
        self.t_arena, self.t_off = t_arena, np.asarray(t_off, dtype=np.int64)
        self.t_len = np.diff(self.t_off)
        self.mean_len = float((self.t_len[order] * weights).sum() / weights.sum())
        self.rng = np.random.Generator(np.random.PCG64(seed + 1000))
        self.seed, self.n_types = seed, n

    @classmethod
    def train5k(cls, seed: int, zipf_c: int = 200_000):
        from subword_tokenizers_b200 import packing as P
        arena, off = P.pack_words(train5k_types())
        return cls(arena, off.astype(np.int64), zipf_c, seed)

    def draws(self, n_words: int):
        """Yields int32 chunks of type ids (deterministic sequence)."""
        left = n_words
        while left > 0:
            k = min(left, DRAW_CHUNK)
            full = self.rng.integers(0, len(self.lut), size=DRAW_CHUNK, dtype=np.int64)   # always a full chunk: prefix-stable
            yield self.lut[full[:k]]
            left -= k

    def host_sample(self, n_words: int):
        """First n_words of the stream as (arena u8, offsets u64) on the host."""
        draw = np.concatenate(list(self.draws(n_words)))
        lens = self.t_len[draw]
        off = np.zeros(n_words + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        idx = np.repeat(self.t_off[:-1][draw] - off[:-1], lens) + np.arange(off[-1])
        return self.t_arena[idx], off.astype(np.uint64)

    def device_stream(self, target_bytes: int, dev):
        """The stream resident on `dev`: (arena u8 tensor, offsets int32-view tensor (u32), n_words, host offsets u32)."""
        import torch
        n_words = int(target_bytes / self.mean_len)
        d_tlen = torch.from_numpy(self.t_len).to(dev)
        d_toff = torch.from_numpy(self.t_off[:-1].copy()).to(dev)
        d_tarena = torch.from_numpy(self.t_arena).to(dev)
        d_draw = torch.empty(n_words, dtype=torch.int32, device=dev)
        p = 0
        for chunk in self.draws(n_words):
            d_draw[p:p + len(chunk)] = torch.from_numpy(chunk).to(dev)
            p += len(chunk)
        lens = d_tlen[d_draw.long()]
        off = torch.zeros(n_words + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=off[1:])
        total = int(off[-1].item())
        assert total < (1 << 32) - 64
        arena = torch.empty(total + 64, dtype=torch.uint8, device=dev)[:total]
        step = 1 << 23
        for a in range(0, n_words, step):
            b = min(n_words, a + step)
            l = lens[a:b]
            o = off[a:b] - off[a]
            nb = int((off[b] - off[a]).item())
            widx = torch.repeat_interleave(torch.arange(b - a, device=dev), l, output_size=nb)
            src = d_toff[d_draw[a:b].long()][widx] + (torch.arange(nb, device=dev) - o[widx])
            arena[int(off[a].item()):int(off[b].item())] = d_tarena[src]
        off32 = off.to(torch.int64).cpu().numpy().astype(np.uint32)
        d_off = torch.from_numpy(off32.view(np.int32)).to(dev)
        return arena, d_off, n_words, off32
