"""ctypes front end of the CPU oracle (oracle/swt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and --impl reference
legs.  The product package (subword_tokenizers_b200) never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib: Optional[ctypes.CDLL] = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "swt_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_bpe_train.restype = ctypes.c_int64
        L.oracle_bpe_train.argtypes = [_u32p, _u64p, _i64p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64,
                                       ctypes.c_int64, _u32p, _u32p, _u32p, _i64p, _i64p]
        L.oracle_bpe_encode.restype = ctypes.c_int64
        L.oracle_bpe_encode.argtypes = [_u32p, _u32p, _u32p, ctypes.c_uint32, _u32p, _u32p, ctypes.c_uint32,
                                        _u8p, _u64p, ctypes.c_uint64, _u32p, ctypes.c_uint64, _u64p, ctypes.c_int]
        L.oracle_wp_build.restype = ctypes.c_void_p
        L.oracle_wp_build.argtypes = [_u32p, _u64p, ctypes.c_uint32, _u8p]
        L.oracle_wp_free.restype = None
        L.oracle_wp_free.argtypes = [ctypes.c_void_p]
        L.oracle_wp_stats.restype = ctypes.c_uint64
        L.oracle_wp_stats.argtypes = [ctypes.c_void_p, _u64p, _u64p, _u64p]
        L.oracle_wp_encode.restype = ctypes.c_int64
        L.oracle_wp_encode.argtypes = [ctypes.c_void_p, _u8p, _u64p, ctypes.c_uint64, _u8p, _u32p, ctypes.c_uint64,
                                       _u64p, _u64p]
        L.oracle_wp_naive_encode.restype = ctypes.c_int64
        L.oracle_wp_naive_encode.argtypes = [ctypes.c_void_p, _u8p, _u64p, ctypes.c_uint64, _u32p, ctypes.c_uint64, _u64p]
        L.oracle_wp_train.restype = ctypes.c_int64
        L.oracle_wp_train.argtypes = [_u32p, _u64p, _i64p, ctypes.c_uint64, _u32p, _u64p, ctypes.c_uint32,
                                      ctypes.c_int64, ctypes.c_int64, _u32p, _u32p, _u32p, _i64p]
        _lib = L
    return _lib


def _p(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def bpe_train(syms, off, freq, n_alpha: int, max_vocab: int, max_merges: Optional[int] = None):
    """-> (left, right, new, count) arrays and the final vocab size."""
    syms, off, freq = _c(syms, np.uint32), _c(off, np.uint64), _c(freq, np.int64)
    n_types = len(off) - 1
    if max_merges is None:
        max_merges = max(16, int(len(syms)) + 16)
    left = np.zeros(max_merges, dtype=np.uint32)
    right = np.zeros(max_merges, dtype=np.uint32)
    new = np.zeros(max_merges, dtype=np.uint32)
    cnt = np.zeros(max_merges, dtype=np.int64)
    vs = ctypes.c_int64(0)
    n = lib().oracle_bpe_train(_p(syms, _u32p), _p(off, _u64p), _p(freq, _i64p), n_types, n_alpha, max_vocab,
                               max_merges, _p(left, _u32p), _p(right, _u32p), _p(new, _u32p), _p(cnt, _i64p),
                               ctypes.byref(vs))
    if n < 0:
        raise RuntimeError("oracle_bpe_train: max_merges too small")
    return left[:n].copy(), right[:n].copy(), new[:n].copy(), cnt[:n].copy(), int(vs.value)


def bpe_encode(tables, arena, off, naive: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """tables: packing.BpeTables.  -> (token ids u32, token offsets u64[n_words+1])."""
    arena, off = _c(arena, np.uint8), _c(off, np.uint64)
    n_words = len(off) - 1
    cap = int(len(arena)) + n_words + 1
    out = np.zeros(cap, dtype=np.uint32)
    tok_off = np.zeros(n_words + 1, dtype=np.uint64)
    left, right, new = _c(tables.left, np.uint32), _c(tables.right, np.uint32), _c(tables.new, np.uint32)
    ccp, cid = _c(tables.char_cp, np.uint32), _c(tables.char_id, np.uint32)
    n = lib().oracle_bpe_encode(_p(left, _u32p), _p(right, _u32p), _p(new, _u32p), len(left), _p(ccp, _u32p),
                                _p(cid, _u32p), len(ccp), _p(arena, _u8p), _p(off, _u64p), n_words,
                                _p(out, _u32p), cap, _p(tok_off, _u64p), 1 if naive else 0)
    if n < 0:
        raise RuntimeError("oracle_bpe_encode: output capacity")
    return out[:n].copy(), tok_off


class WpTrie:
    def __init__(self, tables, alnum_bitmap: np.ndarray):
        """tables: packing.WpTables."""
        cps, off = _c(tables.cps, np.uint32), _c(tables.off, np.uint64)
        self._bm = _c(alnum_bitmap, np.uint8)
        self.n_vocab = tables.n_vocab
        self._h = lib().oracle_wp_build(_p(cps, _u32p), _p(off, _u64p), tables.n_vocab, _p(self._bm, _u8p))

    def stats(self):
        e, p, r = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        n = lib().oracle_wp_stats(self._h, ctypes.byref(e), ctypes.byref(p), ctypes.byref(r))
        return {"nodes": int(n), "edges": int(e.value), "pops": int(p.value), "root_p_links": int(r.value)}

    def encode(self, arena, off, space_bitmap) -> Tuple[np.ndarray, np.ndarray, int]:
        """FastWP.tokenize per word -> (ids, tok offsets, h6_events)."""
        arena, off, sb = _c(arena, np.uint8), _c(off, np.uint64), _c(space_bitmap, np.uint8)
        n_words = len(off) - 1
        cap = int(len(arena)) + 8 * n_words + 8
        out = np.zeros(cap, dtype=np.uint32)
        tok_off = np.zeros(n_words + 1, dtype=np.uint64)
        h6 = ctypes.c_uint64(0)
        n = lib().oracle_wp_encode(self._h, _p(arena, _u8p), _p(off, _u64p), n_words, _p(sb, _u8p), _p(out, _u32p),
                                   cap, _p(tok_off, _u64p), ctypes.byref(h6))
        if n < 0:
            raise RuntimeError("oracle_wp_encode: output capacity")
        return out[:n].copy(), tok_off, int(h6.value)

    def naive_encode(self, arena, off) -> Tuple[np.ndarray, np.ndarray]:
        arena, off = _c(arena, np.uint8), _c(off, np.uint64)
        n_words = len(off) - 1
        cap = int(len(arena)) + 8 * n_words + 8
        out = np.zeros(cap, dtype=np.uint32)
        tok_off = np.zeros(n_words + 1, dtype=np.uint64)
        n = lib().oracle_wp_naive_encode(self._h, _p(arena, _u8p), _p(off, _u64p), n_words, _p(out, _u32p), cap,
                                         _p(tok_off, _u64p))
        if n < 0:
            raise RuntimeError("oracle_wp_naive_encode: output capacity")
        return out[:n].copy(), tok_off

    def close(self):
        if self._h:
            lib().oracle_wp_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def wp_train(syms, off, freq, init_cps, init_off, max_vocab: int, max_merges: Optional[int] = None):
    syms, off, freq = _c(syms, np.uint32), _c(off, np.uint64), _c(freq, np.int64)
    init_cps, init_off = _c(init_cps, np.uint32), _c(init_off, np.uint64)
    n_types, n_init = len(off) - 1, len(init_off) - 1
    if max_merges is None:
        max_merges = max(16, int(len(syms)) + 16)
    left = np.zeros(max_merges, dtype=np.uint32)
    right = np.zeros(max_merges, dtype=np.uint32)
    new = np.zeros(max_merges, dtype=np.uint32)
    vs = ctypes.c_int64(0)
    n = lib().oracle_wp_train(_p(syms, _u32p), _p(off, _u64p), _p(freq, _i64p), n_types, _p(init_cps, _u32p),
                              _p(init_off, _u64p), n_init, max_vocab, max_merges, _p(left, _u32p), _p(right, _u32p),
                              _p(new, _u32p), ctypes.byref(vs))
    if n < 0:
        raise RuntimeError("oracle_wp_train failed (%d)" % n)
    return left[:n].copy(), right[:n].copy(), new[:n].copy(), int(vs.value)
