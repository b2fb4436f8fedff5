"""ctypes binding of libswt.so (include/swt.h).  There is no CPU fallback: if the CUDA library is
missing or no device is usable, every compute call raises."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWT_LIB_PATH") or os.path.join(_HERE, "libswt.so")      # SWT_LIB_PATH: kernel-variant experiments

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p

SWT_OK = 0
SHORT_WORD_BYTES = 32
TRAIN_BPE, TRAIN_WP = 0, 1
PRETOK_PYTHON_SPLIT, PRETOK_BERT = 0, 1


class SwtError(RuntimeError):
    pass


class TrainConfig(ctypes.Structure):
    _fields_ = [
        ("n_types_local", ctypes.c_uint64), ("n_slots_local", ctypes.c_uint64), ("slot_base", ctypes.c_uint64),
        ("n_alpha", ctypes.c_uint32), ("max_vocab", ctypes.c_int64), ("initial_vocab", ctypes.c_int64),
        ("max_word_len", ctypes.c_uint32), ("record_cap", ctypes.c_uint32), ("world_size", ctypes.c_uint32),
        ("rank", ctypes.c_uint32), ("table_cap", ctypes.c_uint64), ("mode", ctypes.c_uint32),
    ]


class TrainState(ctypes.Structure):
    _fields_ = [
        ("halt", ctypes.c_uint32), ("n_recorded", ctypes.c_uint32), ("n_merges_total", ctypes.c_uint64),
        ("vocab_size", ctypes.c_int64), ("n_symbols", ctypes.c_uint64), ("n_table_entries", ctypes.c_uint64),
        ("table_cap", ctypes.c_uint64), ("n_live_slots", ctypes.c_uint64), ("n_tie_steps", ctypes.c_uint64),
        ("n_tie_listed", ctypes.c_uint64), ("n_peer_barriers", ctypes.c_uint64), ("peer_wait_cycles", ctypes.c_uint64),
        ("peer_kernel_cycles", ctypes.c_uint64 * 3),
    ]


# name -> (restype, argtypes); also the list the CPU test checks against include/swt.h
SIGNATURES = {
    "swt_abi_version": (ctypes.c_int, []),
    "swt_last_error": (ctypes.c_char_p, []),
    "swt_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "swt_tune": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "swt_bpe_table_create": (ctypes.c_int, [c_u32p, c_u32p, c_u32p, ctypes.c_uint32, c_u32p, c_u32p, ctypes.c_uint32,
                                            ctypes.c_int, ctypes.POINTER(c_vp)]),
    "swt_bpe_table_destroy": (None, [c_vp]),
    "swt_encode_workspace_bytes": (ctypes.c_size_t, [ctypes.c_uint32, ctypes.c_uint64]),
    "swt_bpe_encode": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint64, c_vp, ctypes.c_uint64, c_vp,
                                      c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "swt_bpe_encode_naive": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint64, c_vp, ctypes.c_uint64, c_vp,
                                            c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "swt_wp_encode_naive": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint64, c_vp, ctypes.c_uint64, c_vp,
                                           c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "swt_wp_trie_create": (ctypes.c_int, [c_u32p, c_u64p, ctypes.c_uint32, c_u8p, c_u32p, ctypes.c_uint32, ctypes.c_int,
                                          ctypes.POINTER(c_vp)]),
    "swt_wp_trie_destroy": (None, [c_vp]),
    "swt_wp_trie_stats": (ctypes.c_int, [c_vp, c_u64p, c_u64p, c_u64p, c_u64p]),
    "swt_wp_encode": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint64, c_vp, ctypes.c_uint64, c_vp,
                                     c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "swt_pretok_create": (ctypes.c_int, [c_u32p, ctypes.c_uint32, c_u32p, ctypes.c_uint32, c_u8p, c_u8p, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(c_vp)]),
    "swt_pretok_destroy": (None, [c_vp]),
    "swt_pretok_workspace_bytes": (ctypes.c_size_t, [ctypes.c_uint64]),
    "swt_pretok_count": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "swt_pretok_write": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp, ctypes.c_size_t, c_vp, ctypes.c_uint64, c_vp, c_vp,
                                        ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, c_vp, c_vp]),
    "swt_types_workspace_bytes": (ctypes.c_size_t, [ctypes.c_uint64, ctypes.c_uint64]),
    "swt_types_count": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint64, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "swt_types_write": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint32, ctypes.c_uint64, c_vp, ctypes.c_size_t, ctypes.c_uint32, c_vp, c_vp,
                                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "swt_types_symbols": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "swt_types_map_symbols": (ctypes.c_int, [c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint32, c_vp, c_vp, ctypes.c_uint32, c_vp]),
    "swt_pipeline_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(c_vp)]),
    "swt_pipeline_destroy": (None, [c_vp]),
    "swt_encode_host": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint64, c_vp,
                                       c_u64p, c_u64p]),
    "swt_encode_host16": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_uint64, c_vp, ctypes.c_uint64, c_vp,
                                         c_u64p, c_u64p]),
    "swt_tokenize_text_host": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_uint64, c_vp, ctypes.c_int, ctypes.c_uint64,
                                              c_u64p, c_u64p, c_u64p]),
    "swt_small_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "swt_small_destroy": (None, [c_vp]),
    "swt_small_max_bytes": (ctypes.c_uint32, []),
    "swt_tokenize_small": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_uint32, ctypes.POINTER(c_vp),
                                          c_u32p, c_u32p, c_u32p]),
    "swt_small_bind": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_u32p]),
    "swt_small_unbind": (None, [c_vp, ctypes.c_uint32]),
    "swt_small_output": (c_vp, [c_vp]),
    "swt_tokenize_small_bound": (ctypes.c_int, [c_vp, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32]),
    "swt_host_alloc": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_size_t]),
    "swt_host_free": (None, [c_vp]),
    "swt_bpe_train_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(TrainConfig)]),
    "swt_bpe_train_create": (ctypes.c_int, [ctypes.POINTER(TrainConfig), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp,
                                            ctypes.POINTER(c_vp)]),
    "swt_bpe_train_destroy": (None, [c_vp]),
    "swt_bpe_train_buffers": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), c_u64p, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                             ctypes.POINTER(c_vp), c_u64p]),
    "swt_bpe_train_count_local": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_build_table": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_export_pairs": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp, c_vp]),
    "swt_bpe_train_import_pairs": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp]),
    "swt_bpe_train_select": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_merge": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_update": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_steps": (ctypes.c_int, [c_vp, ctypes.c_uint32, c_vp]),
    "swt_bpe_train_peer_bytes": (ctypes.c_size_t, [ctypes.POINTER(TrainConfig)]),
    "swt_bpe_train_set_peers": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.c_uint32]),
    "swt_bpe_train_exchange_candidates": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_exchange_deltas": (ctypes.c_int, [c_vp, c_vp]),
    "swt_bpe_train_exchange_probe": (ctypes.c_int, [c_vp, ctypes.c_uint32, c_vp]),
    "swt_bpe_train_read": (ctypes.c_int, [c_vp, c_u32p, c_u32p, c_u32p, c_i64p, ctypes.POINTER(TrainState), c_vp]),
    "swt_bpe_train_table_bytes": (ctypes.c_size_t, [ctypes.c_uint64]),
    "swt_bpe_train_grow_table": (ctypes.c_int, [c_vp, c_vp, ctypes.c_uint64, c_vp]),
    "swt_bpe_train_maintain": (ctypes.c_int, [c_vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int), c_vp]),
    "swt_bpe_train_read_corpus": (ctypes.c_int, [c_vp, c_u32p, c_u32p, c_vp]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Loads libswt.so and declares every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SwtError(
                "libswt.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`python -m subword_tokenizers_b200.build`; there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)            # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if lib.swt_abi_version() != 1:
            raise SwtError("libswt.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != SWT_OK:
        msg = load().swt_last_error()
        raise SwtError("%s failed (status %d): %s" % (what or "libswt call", rc, msg.decode("utf-8", "replace") if msg else ""))


def require_cuda() -> None:
    """Fail loudly when no CUDA device is usable (the product has no CPU path)."""
    import torch
    if not torch.cuda.is_available():
        raise SwtError("no CUDA device available: subword_tokenizers_b200 computes only on the GPU (no CPU fallback)")
    n = ctypes.c_int(0)
    check(load().swt_device_count(ctypes.byref(n)), "swt_device_count")
    if n.value < 1:
        raise SwtError("libswt sees no CUDA device")
