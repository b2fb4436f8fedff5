"""Device-side objects: rank table / trie handles, encode calls and the BPE trainer loop.

PyTorch is used only to own device buffers and streams (and ``torch.distributed`` for the two small
collectives of multi-GPU training); every computation is a call into libswt.so.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import SHORT_WORD_BYTES, SwtError, TrainConfig, TrainState, check, c_u8p, c_u32p, c_u64p, c_i64p, c_vp
from . import packing as P


def _np_ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def tune(name: str, value: int) -> None:
    """Process-wide experiment knob of libswt (include/swt.h, swt_tune)."""
    check(_lib.load().swt_tune(name.encode(), int(value)), "swt_tune")


def current_device() -> int:
    return torch.cuda.current_device()


_PIPELINES: "dict[Tuple[int, int], c_vp]" = {}       # (device, batch_bytes) -> swt_pipeline*, most recently used last


def _pipeline_for(device: int, batch_bytes: int):
    """Pipelines hold no table state, so one per (device, batch size) serves every encoder; at most three sizes are kept."""
    lib = _lib.load()
    key = (device, batch_bytes)
    h = _PIPELINES.pop(key, None)
    if h is None:
        while len(_PIPELINES) >= 3:
            lib.swt_pipeline_destroy(_PIPELINES.pop(next(iter(_PIPELINES))))
        h = c_vp(None)
        check(lib.swt_pipeline_create(device, batch_bytes, ctypes.byref(h)), "swt_pipeline_create")
    _PIPELINES[key] = h
    return h


class SmallCall:
    """Per-device state of the small-call path (swt_small_*): pinned mapped in/out buffers, device scratch, a stream.  One per process
    and device, shared by every encoder."""

    _instances = {}

    @classmethod
    def get(cls, device: Optional[int] = None) -> "SmallCall":
        device = current_device() if device is None else device
        inst = cls._instances.get(device)
        if inst is None:
            inst = cls._instances[device] = cls(device)
        return inst

    def __init__(self, device: int):
        lib = _lib.load()
        self.lib = lib
        self.handle = c_vp(None)
        check(lib.swt_small_create(device, ctypes.byref(self.handle)), "swt_small_create")
        self.max_bytes = int(lib.swt_small_max_bytes())
        self._fn = lib.swt_tokenize_small_bound
        cap = (self.max_bytes * 3 // 2 + 64) + self.max_bytes + 1
        out = lib.swt_small_output(self.handle)
        # numpy view of the pinned output buffer (fixed address): header words 0..7, then the ids
        self._out = np.ctypeslib.as_array(ctypes.cast(out, c_u32p), shape=(cap + 8,))
        self._ids = self._out[8:]

    def bind(self, pretok_handle, which: int, table_handle, naive: bool):
        """Binds a (pre-tokenizer, table) pair once; tokenize() then passes the text only.  -> binding index, or the argument tuple of
        the unbound call when all bindings are in use."""
        k = ctypes.c_uint32(0)
        if self.lib.swt_small_bind(self.handle, pretok_handle, which, table_handle, 1 if naive else 0, ctypes.byref(k)) != 0:
            return (pretok_handle, which, table_handle, 1 if naive else 0)
        return int(k.value)

    def unbind(self, binding) -> None:
        if isinstance(binding, int):
            self.lib.swt_small_unbind(self.handle, binding)

    def tokenize(self, binding: int, data: bytes) -> np.ndarray:
        """-> a COPY of the token ids (u32) of `data` (raw UTF-8 text, at most max_bytes)."""
        if isinstance(binding, int):
            rc = self._fn(self.handle, binding, data, len(data))
        else:
            ids, nt = c_vp(None), ctypes.c_uint32(0)
            rc = self.lib.swt_tokenize_small(self.handle, binding[0], binding[1], binding[2], binding[3], data, len(data), ctypes.byref(ids),
                                             ctypes.byref(nt), None, None)
        if rc:
            check(rc, "swt_tokenize_small")
        return self._ids[:int(self._out[1])].copy()


class _Encoder:
    """Shared encode plumbing: words -> arena on device -> libswt encode -> token ids (+ offsets)."""

    _which = -1

    naive = False          # True: the Naive* encoder of the same table (swt_*_encode_naive)

    def __init__(self):
        self._handle = c_vp(None)
        self._small_ctx = None

    @property
    def _pretok_mode(self) -> int:
        # FastWP works on whitespace chunks; the three other encoders on BERT pre-tokenized words
        return _lib.PRETOK_PYTHON_SPLIT if (self._which == 1 and not self.naive) else _lib.PRETOK_BERT

    # -- device-resident call ---------------------------------------------------------------------------
    def encode_device(self, d_arena: torch.Tensor, d_off: torch.Tensor, n_words: int, long_word_bytes: int,
                      out_cap: Optional[int] = None, want_offsets: bool = True):
        """d_arena: uint8 CUDA tensor, d_off: int32/uint32-compatible CUDA tensor of n_words+1 offsets.
        Returns (d_ids[out_cap] int32 tensor, d_tok_off or None, d_status); asynchronous."""
        lib = _lib.load()
        dev = d_arena.device
        if out_cap is None:
            out_cap = int(d_arena.numel()) + n_words + 16
        ws_bytes = lib.swt_encode_workspace_bytes(n_words, long_word_bytes)
        d_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        d_ids = torch.empty(max(out_cap, 1), dtype=torch.int32, device=dev)
        d_tok_off = torch.empty(n_words + 1, dtype=torch.int32, device=dev) if want_offsets else None
        d_status = torch.empty(8, dtype=torch.int32, device=dev)
        self.encode_into(d_arena, d_off, n_words, long_word_bytes, d_ids, out_cap, d_tok_off, d_ws, d_status)
        return d_ids, d_tok_off, d_status

    def encode_into(self, d_arena, d_off, n_words, long_word_bytes, d_ids, out_cap, d_tok_off, d_ws, d_status):
        lib = _lib.load()
        if self.naive:
            fn = lib.swt_bpe_encode_naive if self._which == 0 else lib.swt_wp_encode_naive
        else:
            fn = lib.swt_bpe_encode if self._which == 0 else lib.swt_wp_encode
        check(fn(self._handle, d_arena.data_ptr(), d_off.data_ptr(), n_words, long_word_bytes, d_ids.data_ptr(), out_cap,
                 d_tok_off.data_ptr() if d_tok_off is not None else None, d_ws.data_ptr(), d_ws.numel(),
                 d_status.data_ptr(), _stream_ptr()), "swt_encode")

    @staticmethod
    def check_status(d_status: torch.Tensor) -> Tuple[int, int]:
        """Synchronises; returns (n_tokens, h6_events) or raises."""
        st = d_status.cpu().numpy().astype(np.uint32)
        if st[0] != 0:
            raise SwtError("encode kernel status %d" % int(st[0]))
        return int(st[1]) | (int(st[3]) << 32), int(st[2])

    # -- convenience: Python words in, numpy out ---------------------------------------------------------------
    def encode_words(self, words: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, int]:
        """-> (token ids u32, token offsets u32[n_words+1], h6_events)."""
        _lib.require_cuda()
        arena, off = P.pack_words(words, offset_dtype=np.uint64)
        if len(arena) >= (1 << 32) - 64:
            raise SwtError("arena >= 4 GiB: split the batch")
        return self.encode_packed(arena, off.astype(np.uint32))

    def encode_packed(self, arena: np.ndarray, off: np.ndarray) -> Tuple[np.ndarray, np.ndarray, int]:
        n_words = len(off) - 1
        lens = np.diff(off.astype(np.int64))
        long_bytes = int(lens[lens > SHORT_WORD_BYTES].sum()) if n_words else 0
        dev = torch.device("cuda", current_device())
        d_arena = torch.from_numpy(arena).to(dev) if len(arena) else torch.zeros(1, dtype=torch.uint8, device=dev)
        d_off = torch.from_numpy(off.view(np.int32)).to(dev)
        d_ids, d_tok_off, d_status = self.encode_device(d_arena, d_off, n_words, long_bytes)
        n_tok, h6 = self.check_status(d_status)
        ids = d_ids[:n_tok].cpu().numpy().view(np.uint32)
        tok_off = d_tok_off.cpu().numpy().view(np.uint32)
        return ids, tok_off, h6

    # -- host-buffer pipeline (the e2e path) --------------------------------------------------------------------
    def pipeline(self, batch_bytes: int = 64 << 20):
        """The staging pipeline (device buffers + streams) for this batch size; shared by all encoders of the process."""
        return _pipeline_for(current_device(), batch_bytes)

    def encode_host(self, h_arena: torch.Tensor, h_off: torch.Tensor, h_out_ids: torch.Tensor,
                    h_out_tok_off: Optional[torch.Tensor], batch_bytes: int = 64 << 20) -> Tuple[int, int]:
        """Host (ideally pinned) tensors in and out, through swt_encode_host (int32/uint32 h_out_ids) or
        swt_encode_host16 (int16/uint16 h_out_ids; raises when an id does not fit). -> (n_tokens, h6_events)."""
        lib = _lib.load()
        n_words = h_off.numel() - 1
        nt, h6 = ctypes.c_uint64(0), ctypes.c_uint64(0)
        fn = lib.swt_encode_host16 if h_out_ids.element_size() == 2 else lib.swt_encode_host
        check(fn(self.pipeline(batch_bytes), self._which, self._handle, h_arena.data_ptr(), h_off.data_ptr(),
                 n_words, h_out_ids.data_ptr(), h_out_ids.numel(),
                 h_out_tok_off.data_ptr() if h_out_tok_off is not None else None,
                 ctypes.byref(nt), ctypes.byref(h6)), fn.__name__)
        return int(nt.value), int(h6.value)

    def tokenize_host(self, h_text, n_bytes: int, h_out_ids, has_sigma: bool = True, batch_bytes: int = 64 << 20) -> Tuple[int, int, int]:
        """Raw UTF-8 text in a host buffer (torch tensor, ideally pinned, or numpy array) -> flat token ids in h_out_ids
        (16- or 32-bit elements), through swt_tokenize_text_host. -> (n_tokens, n_words, h6_events)."""
        lib = _lib.load()
        if self.naive:
            raise SwtError("the host-buffer pipeline serves the Fast encoders only")
        pt = Pretokenizer.get(mode=self._pretok_mode)
        if has_sigma and not pt._with_sigma:
            pt._create(True)
        is_np = isinstance(h_text, np.ndarray)
        text_ptr = h_text.ctypes.data if is_np else h_text.data_ptr()
        out_ptr = h_out_ids.ctypes.data if isinstance(h_out_ids, np.ndarray) else h_out_ids.data_ptr()
        out_n = h_out_ids.size if isinstance(h_out_ids, np.ndarray) else h_out_ids.numel()
        elem = h_out_ids.itemsize if isinstance(h_out_ids, np.ndarray) else h_out_ids.element_size()
        nt, nw, h6 = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        check(lib.swt_tokenize_text_host(self.pipeline(batch_bytes), pt._handle, self._which, self._handle, text_ptr, n_bytes,
                                         out_ptr, 1 if elem == 2 else 0, out_n, ctypes.byref(nt), ctypes.byref(nw), ctypes.byref(h6)),
              "swt_tokenize_text_host")
        return int(nt.value), int(nw.value), int(h6.value)

    BATCH_TEXT_BYTES = 1 << 30            # texts are concatenated up to this many bytes per device pass

    def encode_texts(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """tokenize() of many texts in one pass over their concatenation (the harness batch entry point, SURVEY.md §8f row 4;
        several passes when the texts exceed BATCH_TEXT_BYTES together).
        -> (token ids u32 of all texts back to back, int64 offsets[len(texts) + 1] of each text's tokens)."""
        enc = [P.encode_utf8(t) for t in texts]
        if sum(len(b) + 1 for b in enc) > self.BATCH_TEXT_BYTES and len(enc) > 1:
            ids_parts, cuts, base, k = [], [np.zeros(1, np.int64)], 0, 0
            while k < len(enc):
                k1, size = k, 0
                while k1 < len(enc) and (k1 == k or size + len(enc[k1]) + 1 <= self.BATCH_TEXT_BYTES):
                    size += len(enc[k1]) + 1
                    k1 += 1
                ids, cut = self._encode_texts_once(enc[k:k1])
                ids_parts.append(ids); cuts.append(cut[1:] + base)
                base += len(ids); k = k1
            return np.concatenate(ids_parts), np.concatenate(cuts)
        return self._encode_texts_once(enc)

    def _encode_texts_once(self, enc: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
        texts = enc
        bounds = np.zeros(len(texts) + 1, dtype=np.int64)
        np.cumsum([len(b) + 1 for b in enc], out=bounds[1:])                 # +1: the "\n" that separates two texts
        data = b"\n".join(enc)
        if not data:
            return np.zeros(0, np.uint32), np.zeros(len(texts) + 1, np.int64)
        d_arena, d_off, n_words, d_src = Pretokenizer.get(mode=self._pretok_mode).split_text(data, want_src=True)
        if n_words == 0:
            return np.zeros(0, np.uint32), np.zeros(len(texts) + 1, np.int64)
        lens = (d_off[1:n_words + 1] - d_off[:n_words]).long()                    # bytes of the long words size the scratch
        long_bytes = int(lens[lens > SHORT_WORD_BYTES].sum().item())
        d_ids, d_tok_off, d_status = self.encode_device(d_arena, d_off, n_words, long_bytes)
        n_tok, _ = self.check_status(d_status)
        ids = d_ids[:n_tok].cpu().numpy().view(np.uint32)
        tok_off = d_tok_off.cpu().numpy().view(np.uint32).astype(np.int64)
        first_word = np.searchsorted(d_src[:n_words].cpu().numpy().view(np.uint32), bounds, side="left")
        return ids, tok_off[first_word]

    SMALL_TEXT_BYTES = 1 << 20

    def encode_text(self, text: str, return_offsets: bool = False):
        """tokenize() on the device from the raw text: pre-tokenization (lower-casing, splitting) + encode.  Texts up to 1 MiB take
        the single C call (swt_tokenize_text_host, three synchronisations); larger ones stay resident (Pretokenizer + encode_device).
        -> token ids u32 (and the u32 token offsets per word when return_offsets)."""
        data = P.encode_utf8(text)
        if not return_offsets:
            ctx = self._small_ctx
            if ctx is not None and ctx[2] != ctx[1]._handle.value:
                ctx[0].unbind(ctx[3])
                ctx = None
            if ctx is None:
                # (small-call state, pre-tokenizer, its handle, binding) of this encoder's device, set up once (and again when
                # the pre-tokenizer was re-created with the sigma bitmaps)
                sc, pt = SmallCall.get(), Pretokenizer.get(mode=self._pretok_mode)
                ctx = self._small_ctx = (sc, pt, pt._handle.value, sc.bind(pt._handle, self._which, self._handle, self.naive))
            sc = ctx[0]
            if len(data) <= sc.max_bytes:
                # one short text (the per-line pattern of the reference's CLI): ONE single-CTA kernel, zero-copy buffers, no stream sync
                if not ctx[1]._with_sigma and "\u03a3" in text:
                    ctx[1]._create(True)
                    return self.encode_text(text)
                return sc.tokenize(ctx[3], data)
        if not return_offsets and not self.naive and len(data) <= self.SMALL_TEXT_BYTES:
            if not data:
                return np.zeros(0, np.uint32)
            out = np.empty(2 * len(data) + 16, dtype=np.uint32)
            nt, _, _ = self.tokenize_host(np.frombuffer(data, dtype=np.uint8), len(data), out, has_sigma="\u03a3" in text,
                                          batch_bytes=self.SMALL_TEXT_BYTES + 64)
            return out[:nt]
        return self._encode_text_resident(text, return_offsets)

    def close(self):
        ctx, self._small_ctx = self._small_ctx, None
        if ctx is not None:
            ctx[0].unbind(ctx[3])              # the small-call binding points at the table about to be destroyed
        self._destroy()

    def _destroy(self):
        raise NotImplementedError

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pretokenizer:
    """Device pre-tokenizer: raw UTF-8 text -> lower-cased word arena + offsets.  PRETOK_PYTHON_SPLIT = ``text.lower().split()``
    of FastWP.tokenize (reference wordpiece.py:248,266-269); PRETOK_BERT = ``pre_tokenize_str(text.lower())`` of the BPE classes
    (utils.py:26-29, Rust BertPreTokenizer).  One instance per process, device and mode; the sigma bitmaps are uploaded on the
    first text that contains U+03A3."""

    _instances = {}

    @classmethod
    def get(cls, device: Optional[int] = None, mode: int = _lib.PRETOK_PYTHON_SPLIT) -> "Pretokenizer":
        device = current_device() if device is None else device
        if (device, mode) not in cls._instances:
            cls._instances[(device, mode)] = cls(device, mode)
        return cls._instances[(device, mode)]

    def __init__(self, device: int, mode: int = _lib.PRETOK_PYTHON_SPLIT):
        _lib.require_cuda()
        self.device = device
        self.mode = mode
        self.tables = P.PretokTables.get()
        self._handle = c_vp(None)
        self._with_sigma = False
        self._create(False)

    def _create(self, with_sigma: bool):
        lib = _lib.load()
        if self._handle:
            lib.swt_pretok_destroy(self._handle)
            self._handle = c_vp(None)
        t = self.tables
        lower = t.bert_lower_map() if self.mode == _lib.PRETOK_BERT else t.lower_map
        cased = ign = None
        if with_sigma:
            cased, ign = t.sigma_bitmaps()
        check(lib.swt_pretok_create(_np_ptr(lower, c_u32p), len(lower), _np_ptr(t.multi, c_u32p), len(t.multi),
                                    _np_ptr(cased, c_u8p) if with_sigma else None, _np_ptr(ign, c_u8p) if with_sigma else None,
                                    self.mode, self.device, ctypes.byref(self._handle)), "swt_pretok_create")
        self._with_sigma = with_sigma

    def split_device(self, d_text: torch.Tensor, n_bytes: int, has_sigma: bool, want_src: bool = False):
        """d_text: uint8 CUDA tensor holding n_bytes of UTF-8 (numel a multiple of 4). -> (d_arena, d_word_off, n_words), plus the
        int32 tensor of the words' start positions in the text when want_src."""
        lib = _lib.load()
        if has_sigma and not self._with_sigma:
            self._create(True)
        dev = d_text.device
        ws = torch.empty(lib.swt_pretok_workspace_bytes(n_bytes), dtype=torch.uint8, device=dev)
        d_status = torch.empty(8, dtype=torch.int32, device=dev)
        check(lib.swt_pretok_count(self._handle, d_text.data_ptr(), n_bytes, ws.data_ptr(), ws.numel(), d_status.data_ptr(),
                                   _stream_ptr()), "swt_pretok_count")
        st = d_status.cpu().numpy().astype(np.uint32)
        if st[0] != 0:
            raise SwtError("pre-tokenizer status %d" % int(st[0]))
        n_words, n_out = int(st[1]), int(st[2]) | (int(st[3]) << 32)
        d_arena = torch.empty(n_out + 16, dtype=torch.uint8, device=dev)
        d_off = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
        d_src = torch.empty(max(n_words, 1), dtype=torch.int32, device=dev) if want_src else None
        check(lib.swt_pretok_write(self._handle, d_text.data_ptr(), n_bytes, ws.data_ptr(), ws.numel(), d_arena.data_ptr(), n_out,
                                   d_off.data_ptr(), d_src.data_ptr() if want_src else None, n_words + 1, n_words, n_out,
                                   d_status.data_ptr(), _stream_ptr()), "swt_pretok_write")
        return (d_arena, d_off, n_words, d_src) if want_src else (d_arena, d_off, n_words)

    def split_text(self, text, want_src: bool = False):
        """Python str (or its UTF-8 bytes) -> (d_arena, d_word_off, n_words[, d_word_src]) on the current device."""
        data = text if isinstance(text, (bytes, bytearray)) else P.encode_utf8(text)
        n = len(data)
        if n >= (1 << 32) - 256:
            raise SwtError("text >= 4 GiB: split it")
        host = np.zeros((n + 7) // 4 * 4, dtype=np.uint8)
        host[:n] = np.frombuffer(data, dtype=np.uint8)
        d_text = torch.from_numpy(host).to(torch.device("cuda", self.device))
        has_sigma = (b"\xce\xa3" in data) if isinstance(text, (bytes, bytearray)) else ("\u03a3" in text)
        return self.split_device(d_text, n, has_sigma, want_src)


class BpeEncoder(_Encoder):
    """Device rank table of a merge list (reference FastBPE._bpe_ranks, bpe.py:200,257)."""

    _which = 0

    def __init__(self, tables: P.BpeTables, device: Optional[int] = None, naive: bool = False):
        super().__init__()
        self.naive = naive
        _lib.require_cuda()
        lib = _lib.load()
        self.tables = tables
        left, right, new = (np.ascontiguousarray(x, dtype=np.uint32) for x in (tables.left, tables.right, tables.new))
        ccp, cid = np.ascontiguousarray(tables.char_cp, np.uint32), np.ascontiguousarray(tables.char_id, np.uint32)
        check(lib.swt_bpe_table_create(_np_ptr(left, c_u32p), _np_ptr(right, c_u32p), _np_ptr(new, c_u32p), len(left),
                                       _np_ptr(ccp, c_u32p), _np_ptr(cid, c_u32p), len(ccp),
                                       current_device() if device is None else device, ctypes.byref(self._handle)),
              "swt_bpe_table_create")

    def _destroy(self):
        if self._handle:
            _lib.load().swt_bpe_table_destroy(self._handle)
            self._handle = c_vp(None)


def _bpe_encode_text_resident(self, text: str, return_offsets: bool = False):
    """FastBPE.tokenize on the device from the raw text: BERT pre-tokenization (Pretokenizer, PRETOK_BERT) + encode."""
    d_arena, d_off, n_words = Pretokenizer.get(mode=self._pretok_mode).split_text(text)
    if n_words == 0:
        return (np.zeros(0, np.uint32), np.zeros(1, np.uint32)) if return_offsets else np.zeros(0, np.uint32)
    lens = (d_off[1:n_words + 1] - d_off[:n_words]).long()                    # bytes of the long words size the scratch
    long_bytes = int(lens[lens > SHORT_WORD_BYTES].sum().item())
    d_ids, d_tok_off, d_status = self.encode_device(d_arena, d_off, n_words, long_bytes, want_offsets=return_offsets)
    n_tok, _ = self.check_status(d_status)
    ids = d_ids[:n_tok].cpu().numpy().view(np.uint32)
    return (ids, d_tok_off.cpu().numpy().view(np.uint32)) if return_offsets else ids


BpeEncoder._encode_text_resident = _bpe_encode_text_resident


class WpEncoder(_Encoder):
    """Device trie + failure links of a vocabulary (reference WPTrie_E2E, utils.py:66-139)."""

    _which = 1

    def __init__(self, tables: P.WpTables, sharp_special: Sequence[int], device: Optional[int] = None, naive: bool = False):
        super().__init__()
        self.naive = naive
        _lib.require_cuda()
        lib = _lib.load()
        self.tables = tables
        alnum, _ = P.unicode_class_bitmaps()
        cps, off = np.ascontiguousarray(tables.cps, np.uint32), np.ascontiguousarray(tables.off, np.uint64)
        ss = np.ascontiguousarray(list(sharp_special) + [0, 0], dtype=np.uint32)
        if len(sharp_special) > 2:
            raise NotImplementedError("NaiveWP.encode_word('##') longer than 2 tokens is outside the supported domain")
        check(lib.swt_wp_trie_create(_np_ptr(cps, c_u32p), _np_ptr(off, c_u64p), tables.n_vocab, _np_ptr(alnum, c_u8p),
                                     _np_ptr(ss, c_u32p), len(sharp_special),
                                     current_device() if device is None else device, ctypes.byref(self._handle)),
              "swt_wp_trie_create")

    def _encode_text_resident(self, text: str, return_offsets: bool = False):
        d_arena, d_off, n_words = Pretokenizer.get(mode=self._pretok_mode).split_text(text)
        if n_words == 0:
            return (np.zeros(0, np.uint32), np.zeros(1, np.uint32)) if return_offsets else np.zeros(0, np.uint32)
        lens = (d_off[1:n_words + 1] - d_off[:n_words]).long()                    # long chunks are split across a warp (scratch)
        long_bytes = int(lens[lens > SHORT_WORD_BYTES].sum().item())
        d_ids, d_tok_off, d_status = self.encode_device(d_arena, d_off, n_words, long_bytes, want_offsets=return_offsets)
        n_tok, _ = self.check_status(d_status)
        ids = d_ids[:n_tok].cpu().numpy().view(np.uint32)
        return (ids, d_tok_off.cpu().numpy().view(np.uint32)) if return_offsets else ids

    def stats(self):
        n, e, p, r = (ctypes.c_uint64(0) for _ in range(4))
        check(_lib.load().swt_wp_trie_stats(self._handle, ctypes.byref(n), ctypes.byref(e), ctypes.byref(p), ctypes.byref(r)))
        return {"nodes": n.value, "edges": e.value, "pops": p.value, "root_p_links": r.value}

    def _destroy(self):
        if self._handle:
            _lib.load().swt_wp_trie_destroy(self._handle)
            self._handle = c_vp(None)


# ------------------------------------------------------------------------------------------------------------
# word-type table of a corpus, built on the device (pre-tokenization + dedupe in front of the trainers)
# ------------------------------------------------------------------------------------------------------------

def device_train_types(corpus: Sequence[str], wordpiece: bool):
    """Corpus (list of texts) -> packing.TrainTypes / packing.WpTrainTypes with the same content as the host constructors
    on ``pre_tokenize_str(example.lower())`` words, computed on the GPU: BERT pre-tokenization (swt_pretok_*), hash dedupe in
    first-occurrence order with frequencies (swt_types_*), symbol ids.  (WordPiece symbol ids are numbered by code point
    instead of by first occurrence; the trainer's result does not depend on the numbering.)"""
    _lib.require_cuda()
    lib = _lib.load()
    data = b"\n".join(P.encode_utf8(t) for t in corpus)
    dev = torch.device("cuda", current_device())
    n_words = 0
    if data:
        d_arena, d_off, n_words = Pretokenizer.get(mode=_lib.PRETOK_BERT).split_text(data)
    if n_words == 0:                                    # empty corpus: the host constructors give the empty tables
        return P.WpTrainTypes([]) if wordpiece else P.TrainTypes([])
    sp = _stream_ptr()
    d_status = torch.empty(8, dtype=torch.int32, device=dev)
    max_types = max(1024, min(n_words, 1 << 22))
    while True:
        ws = torch.empty(lib.swt_types_workspace_bytes(n_words, max_types), dtype=torch.uint8, device=dev)
        check(lib.swt_types_count(d_arena.data_ptr(), d_off.data_ptr(), n_words, max_types, ws.data_ptr(), ws.numel(),
                                  d_status.data_ptr(), sp), "swt_types_count")
        st = d_status.cpu().numpy().astype(np.uint32)
        if st[0] == 0 and int(st[1]) <= max_types:
            break
        if max_types >= n_words:
            raise SwtError("swt_types_count status %d" % int(st[0]))
        max_types = min(n_words, max_types * 8)
    n_types = int(st[1])
    d_type_word = torch.empty(max(n_types, 1), dtype=torch.int32, device=dev)
    d_freq = torch.empty(max(n_types, 1), dtype=torch.int64, device=dev)
    d_nch = torch.empty(max(n_types, 1), dtype=torch.int32, device=dev)
    d_soff = torch.empty(n_types + 1, dtype=torch.int64, device=dev)
    d_bm1 = torch.empty(0x110000 // 32, dtype=torch.int32, device=dev)
    d_bm2 = torch.empty(0x110000 // 32, dtype=torch.int32, device=dev)
    check(lib.swt_types_write(d_arena.data_ptr(), d_off.data_ptr(), n_words, max_types, ws.data_ptr(), ws.numel(), n_types,
                              d_type_word.data_ptr(), d_freq.data_ptr(), d_nch.data_ptr(), d_soff.data_ptr(), d_bm1.data_ptr(),
                              d_bm2.data_ptr(), d_status.data_ptr(), sp), "swt_types_write")
    st = d_status.cpu().numpy().astype(np.uint32)
    n_syms = int(st[2]) | (int(st[3]) << 32)
    d_syms = torch.empty(max(n_syms, 1), dtype=torch.int32, device=dev)
    check(lib.swt_types_symbols(d_arena.data_ptr(), d_off.data_ptr(), d_type_word.data_ptr(), n_types, d_soff.data_ptr(),
                                d_syms.data_ptr(), d_bm1.data_ptr(), d_bm2.data_ptr(), sp), "swt_types_symbols")
    first = np.flatnonzero(np.unpackbits(d_bm1.cpu().numpy().view(np.uint8), bitorder="little"))
    later = np.flatnonzero(np.unpackbits(d_bm2.cpu().numpy().view(np.uint8), bitorder="little"))
    if wordpiece:
        out = object.__new__(P.WpTrainTypes)
        out.init_syms = [chr(c) for c in first] + ["##" + chr(c) for c in later]
        ids_first, ids_later = (first, np.arange(len(first))), (later, len(first) + np.arange(len(later)))
    else:
        out = object.__new__(P.TrainTypes)
        alpha = np.union1d(first, later)
        out.alphabet = [chr(c) for c in alpha]
        ids_first = ids_later = (alpha, np.arange(len(alpha)))
    n_lut = int(max(first.max() if len(first) else 0, later.max() if len(later) else 0)) + 1
    luts = []
    for cps, ids in (ids_first, ids_later):
        lut = np.full(n_lut, 0xFFFFFFFF, dtype=np.uint32)
        lut[cps] = ids
        luts.append(torch.from_numpy(lut.view(np.int32)).to(dev))
    check(lib.swt_types_map_symbols(d_syms.data_ptr(), n_syms, d_soff.data_ptr(), n_types, luts[0].data_ptr(), luts[1].data_ptr(),
                                    n_lut, sp), "swt_types_map_symbols")
    out.freq = d_freq[:n_types].cpu().numpy()
    out.off = d_soff.cpu().numpy().view(np.uint64)
    out.syms = d_syms[:n_syms].cpu().numpy().view(np.uint32)
    # the type strings are only materialised when somebody asks for them
    h_arena = h_off = None
    type_word = d_type_word[:n_types].cpu().numpy().view(np.uint32)

    def type_strings():
        nonlocal h_arena, h_off
        if h_arena is None:
            h_arena, h_off = d_arena.cpu().numpy(), d_off.cpu().numpy().view(np.uint32)
        return [P.decode_utf8(h_arena[int(h_off[w]):int(h_off[w + 1])].tobytes()) for w in type_word]
    out._type_strings = type_strings
    out._n_types = n_types
    if wordpiece:
        out.init_cps, out.init_off = P.pack_strings_as_cps(out.init_syms)
    return out


# ------------------------------------------------------------------------------------------------------------
# HP-3 trainer
# ------------------------------------------------------------------------------------------------------------

HALT_RUN, HALT_DONE_VOCAB, HALT_DONE_NOPAIRS, HALT_GROW, HALT_RECORD_FULL = 0, 1, 2, 3, 4


def shard_types(off: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous type ranges [t0, t1) per rank, balanced by symbol count (global first-occurrence
    order is preserved inside and across shards, which the tie-break needs)."""
    n_types = len(off) - 1
    total = int(off[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        t = int(np.searchsorted(off, target, side="left"))
        bounds.append(min(max(t, bounds[-1]), n_types))
    bounds.append(n_types)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class CudaTrainEngine:
    """One rank's share of the word table + its replica of the pair table, on the GPU (libswt)."""

    def __init__(self, syms: np.ndarray, off: np.ndarray, freq: np.ndarray, n_alpha: int, max_vocab: int,
                 initial_vocab: int, max_word_len: int, slot_base: int, rank: int, world_size: int,
                 record_cap: int = 4096, table_cap: int = 0, mode: int = 0, init_cps: Optional[np.ndarray] = None,
                 init_off: Optional[np.ndarray] = None):
        """mode 0: BPE (symbols 0..n_alpha-1 are single characters).  mode 1: WordPiece (NaiveWP.train); symbols
        0..n_alpha-1 are the initial symbol strings given by init_cps / init_off."""
        _lib.require_cuda()
        lib = _lib.load()
        self.lib = lib
        self.dev = torch.device("cuda", current_device())
        n_types = len(off) - 1
        self.cfg = TrainConfig(n_types, int(off[-1]) if n_types else 0, slot_base, n_alpha, max_vocab, initial_vocab,
                               max(1, max_word_len), record_cap, world_size, rank, table_cap, mode)
        ws_bytes = lib.swt_bpe_train_workspace_bytes(ctypes.byref(self.cfg))
        self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=self.dev)
        self._keep = [
            torch.from_numpy(np.ascontiguousarray(syms, dtype=np.uint32).view(np.int32)).to(self.dev),
            torch.from_numpy(np.ascontiguousarray(off, dtype=np.uint64).view(np.int64)).to(self.dev),
            torch.from_numpy(np.ascontiguousarray(freq, dtype=np.int64)).to(self.dev),
        ]
        if mode == _lib.TRAIN_WP:
            self._keep.append(torch.from_numpy(np.ascontiguousarray(init_cps, dtype=np.uint32).view(np.int32)).to(self.dev))
            self._keep.append(torch.from_numpy(np.ascontiguousarray(init_off, dtype=np.uint64).view(np.int64)).to(self.dev))
        self.handle = c_vp(None)
        # the trainer runs on its own (non-default) stream: the single-rank loop is replayed as a CUDA graph, and
        # stream capture is not possible on the legacy default stream
        torch.cuda.synchronize()
        self.stream = torch.cuda.Stream(device=self.dev)
        check(lib.swt_bpe_train_create(ctypes.byref(self.cfg), self._keep[0].data_ptr(), self._keep[1].data_ptr(),
                                       self._keep[2].data_ptr(),
                                       self._keep[3].data_ptr() if mode == _lib.TRAIN_WP else None,
                                       self._keep[4].data_ptr() if mode == _lib.TRAIN_WP else None,
                                       self.workspace.data_ptr(), ws_bytes, self._sp(),
                                       ctypes.byref(self.handle)), "swt_bpe_train_create")
        ic, ice, cp, cgp, dp, de = c_vp(), ctypes.c_uint64(), c_vp(), c_vp(), c_vp(), ctypes.c_uint64()
        check(lib.swt_bpe_train_buffers(self.handle, ctypes.byref(ic), ctypes.byref(ice), ctypes.byref(cp), ctypes.byref(cgp),
                                        ctypes.byref(dp), ctypes.byref(de)))
        base = self.workspace.data_ptr()
        w64 = self.workspace.view(torch.int64)           # workspace regions are 256-byte aligned

        def view(ptr, n):
            o = (ptr.value - base) // 8
            return w64[o:o + n]
        self.init_counts = view(ic, ice.value)
        self.cand = view(cp, 2)
        self.cand_gather = view(cgp, 2 * world_size)
        self.delta = view(dp, de.value)
        self.record_cap = record_cap
        self._rec = [np.zeros(record_cap, dtype=np.uint32) for _ in range(3)] + [np.zeros(record_cap, dtype=np.int64)]
        self._tables: List[torch.Tensor] = []
        self.n_types, self.n_slots = n_types, int(off[-1]) if n_types else 0

    def _sp(self) -> int:
        return self.stream.cuda_stream

    exchange_kind = "nccl all_gather + all_reduce per step"

    def setup_peer_exchange(self, group=None) -> bool:
        """Sharded runs on one box: replaces the two NCCL collectives of every step by the trainer's own peer-memory exchange
        (include/swt.h, swt_bpe_train_set_peers).  The exchange buffer is symmetric memory (torch.distributed._symmetric_memory:
        every rank maps every rank's buffer over NVLink); the kernels that push and barrier are libswt's.  -> True when set up."""
        import os
        import torch.distributed as dist
        world = self.cfg.world_size
        if world < 2 or world > 8 or os.environ.get("SWT_NO_PEER_EXCHANGE"):
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = int(self.lib.swt_bpe_train_peer_bytes(ctypes.byref(self.cfg)))
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.dev)
            buf.zero_()
            grp = group if group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(buf, grp)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != world or not all(ptrs):
                return False
            torch.cuda.synchronize()
            dist.barrier(group=group)                      # every buffer is zeroed before anybody pushes into it
            arr = (c_vp * world)(*[c_vp(p) for p in ptrs])
            check(self.lib.swt_bpe_train_set_peers(self.handle, arr, world), "swt_bpe_train_set_peers")
            self._peer = (buf, hdl)                        # keep the mapping alive
            self.peer_exchange = True
            self.exchange_kind = "peer memory: P2P stores + flag barrier inside the step graph (1 kernel per step, 2 on tie steps)"
            return True
        except Exception as e:                              # noqa: BLE001 - any failure keeps the NCCL exchange
            self.peer_error = repr(e)
            return False

    # phases (all asynchronous on the trainer's stream)
    def count_local(self): check(self.lib.swt_bpe_train_count_local(self.handle, self._sp()))
    def build_table(self): check(self.lib.swt_bpe_train_build_table(self.handle, self._sp()))
    def select(self): check(self.lib.swt_bpe_train_select(self.handle, self._sp()))
    def merge(self): check(self.lib.swt_bpe_train_merge(self.handle, self._sp()))
    def update(self): check(self.lib.swt_bpe_train_update(self.handle, self._sp()))
    def steps(self, n: int): check(self.lib.swt_bpe_train_steps(self.handle, n, self._sp()))

    def exchange_initial_pairs(self, world_size: int, group, dist) -> None:
        """Large alphabets (no dense n_alpha^2 array): all-gather the ranks' local (pair, count) lists and add the other
        ranks' entries, so that every replica of the pair table holds the global counts (include/swt.h)."""
        n_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        check(self.lib.swt_bpe_train_export_pairs(self.handle, None, 0, n_dev.data_ptr(), self._sp()))    # count only
        n_local = int(n_dev.item())
        counts = torch.zeros(world_size, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(counts, n_dev, group=group)
        counts = counts.cpu().tolist()
        cap = max(max(counts), 1)
        mine = torch.zeros(2 * cap, dtype=torch.int64, device=self.dev)
        check(self.lib.swt_bpe_train_export_pairs(self.handle, mine.data_ptr(), cap, n_dev.data_ptr(), self._sp()))
        everyone = torch.empty(world_size * 2 * cap, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        rank = self.cfg.rank
        for r in range(world_size):
            if r != rank and counts[r]:
                check(self.lib.swt_bpe_train_import_pairs(self.handle, everyone[2 * cap * r:].data_ptr(), counts[r], self._sp()))
        self.stream.synchronize()
        assert n_local == counts[rank]

    def read(self):
        """Synchronises. -> (state dict, left, right, new, count) for the merges recorded since the last read."""
        st = TrainState()
        check(self.lib.swt_bpe_train_read(self.handle, _np_ptr(self._rec[0], c_u32p), _np_ptr(self._rec[1], c_u32p),
                                          _np_ptr(self._rec[2], c_u32p), _np_ptr(self._rec[3], c_i64p), ctypes.byref(st),
                                          self._sp()), "swt_bpe_train_read")
        n = st.n_recorded
        state = {f[0]: getattr(st, f[0]) for f in TrainState._fields_}
        return state, self._rec[0][:n].copy(), self._rec[1][:n].copy(), self._rec[2][:n].copy(), self._rec[3][:n].copy()

    def maintain(self, n_live_slots: int) -> bool:
        """Filter rebuild / word-table compaction between batches of steps; True when captured graphs must be re-captured."""
        changed = ctypes.c_int(0)
        check(self.lib.swt_bpe_train_maintain(self.handle, int(n_live_slots), ctypes.byref(changed), self._sp()), "swt_bpe_train_maintain")
        return bool(changed.value)

    def grow_table(self, cur_cap: int):
        new_cap = cur_cap * 2
        buf = torch.empty(self.lib.swt_bpe_train_table_bytes(new_cap), dtype=torch.uint8, device=self.dev)
        self._tables.append(buf)                       # keep alive; older tables are released after the rehash ran
        check(self.lib.swt_bpe_train_grow_table(self.handle, buf.data_ptr(), new_cap, self._sp()))
        self.stream.synchronize()
        self._tables = self._tables[-1:]

    def read_corpus(self):
        syms = np.zeros(max(self.n_slots, 1), dtype=np.uint32)
        lens = np.zeros(max(self.n_types, 1), dtype=np.uint32)
        check(self.lib.swt_bpe_train_read_corpus(self.handle, _np_ptr(syms, c_u32p), _np_ptr(lens, c_u32p), self._sp()))
        return syms[:self.n_slots], lens[:self.n_types]

    def close(self):
        if self.handle:
            torch.cuda.synchronize()
            self.lib.swt_bpe_train_destroy(self.handle)
            self.handle = c_vp(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_training_loop(engine, world_size: int = 1, group=None, steps_per_sync: int = 256, progress=None):
    """Drives one rank's engine until the trainer halts; returns (left, right, new, count, state).

    ``engine`` exposes count_local/build_table/select/merge/update/steps/read/grow_table and the exchange
    tensors init_counts / cand / cand_gather / delta.  With world_size > 1 the two collectives per step are
    issued through torch.distributed on those tensors (NCCL on GPUs; the CPU tests drive a numpy engine
    over gloo through this same function).
    """
    import contextlib
    import torch.distributed as dist
    stream = getattr(engine, "stream", None)
    ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
    with ctx:                                   # the collectives must be ordered with the kernels: same stream
        return _training_loop(engine, world_size, group, steps_per_sync, progress, dist)


def _one_step_with_collectives(engine, group, dist):
    engine.select()
    dist.all_gather_into_tensor(engine.cand_gather, engine.cand, group=group)      # 16-byte tie-break candidates
    engine.merge()
    dist.all_reduce(engine.delta, op=dist.ReduceOp.SUM, group=group)               # pair-count deltas L | R | ZZ | M
    engine.update()


STEPS_PER_GRAPH = 16


def _capture_steps(engine, group, dist):
    """Multi-GPU steps are latency-bound by five host calls each; capture a run of steps (kernels + both NCCL
    collectives) into one CUDA graph so that a replay costs one launch.  Returns None when capture is unavailable."""
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=engine.stream, capture_error_mode="thread_local"):
            for _ in range(STEPS_PER_GRAPH):
                _one_step_with_collectives(engine, group, dist)
        return g
    except Exception:                                   # noqa: BLE001 - fall back to eager stepping
        torch.cuda.synchronize()
        return None


def _training_loop(engine, world_size, group, steps_per_sync, progress, dist):
    lefts, rights, news, counts = [], [], [], []
    engine.count_local()
    if world_size > 1:
        if engine.init_counts.numel():
            dist.all_reduce(engine.init_counts, op=dist.ReduceOp.SUM, group=group)
        else:                                   # alphabet > 4096 symbols: sparse exchange of the initial counts
            engine.exchange_initial_pairs(world_size, group, dist)
    engine.build_table()
    graph = None
    # sharded on GPUs of one box: the per-step exchange runs over peer memory inside the step graph (no collective call per step)
    peer = world_size > 1 and hasattr(engine, "setup_peer_exchange") and engine.setup_peer_exchange(group)
    use_graph = world_size > 1 and not peer and getattr(engine, "stream", None) is not None and steps_per_sync >= STEPS_PER_GRAPH
    if use_graph:
        _one_step_with_collectives(engine, group, dist)             # eager once: communicators and kernels are warm
        engine.stream.synchronize()
        graph = _capture_steps(engine, group, dist)
    while True:
        if world_size == 1 or peer:
            engine.steps(steps_per_sync)
        else:
            done = 0
            if graph is not None:
                for _ in range(steps_per_sync // STEPS_PER_GRAPH):
                    graph.replay()
                    done += STEPS_PER_GRAPH
            for _ in range(steps_per_sync - done):
                _one_step_with_collectives(engine, group, dist)
        state, l, r, n, c = engine.read()
        lefts.append(l); rights.append(r); news.append(n); counts.append(c)
        if progress is not None and len(l):
            progress(len(l))
        halt = state["halt"]
        if halt in (HALT_DONE_VOCAB, HALT_DONE_NOPAIRS):
            break
        recapture = False
        if halt == HALT_GROW:
            engine.grow_table(state["table_cap"])
            recapture = True                                         # the captured kernels hold the old table pointer
        elif halt >= 16:
            raise SwtError("BPE trainer halted with device error %d" % halt)
        if hasattr(engine, "maintain"):
            recapture = engine.maintain(state["n_live_slots"]) or recapture
        if recapture and graph is not None:
            graph = _capture_steps(engine, group, dist)
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt)
    return cat(lefts, np.uint32), cat(rights, np.uint32), cat(news, np.uint32), cat(counts, np.int64), state
