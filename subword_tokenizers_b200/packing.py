"""Host-side packing: Python strings <-> the flat integer layouts the C ABI consumes.

Nothing here computes a tokenization; it only changes representation:

* words            -> UTF-8 byte arena + offsets          (north-star subsystem 1)
* merges_list      -> (left id, right id, merged id) rank table with one id per DISTINCT
                      string (reference keys are ``(str, str)``, bpe.py:200)
* vocab (set[str]) -> sorted token list as code points    (reference trie input, utils.py:83-84)
* word types       -> alphabet-id sequences + frequencies (reference bpe.py:73-81)
* Python's ``str.isalnum`` / ``str.isspace`` as code-point bitmaps (reference
  wordpiece.py:285-288 and utils.py:137 use the Python predicates)
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

UNICODE_LIMIT = 0x110000

# token id conventions of the C ABI (include/swt.h)
BPE_UNKNOWN_CP = 0x40000000      # symbol for a code point no merge mentions
BPE_EMPTY_TOKEN = 0xFFFFFFFE     # FastBPE.encode_word("") == [""]  (bpe.py:207-208)

_class_cache: Dict[str, np.ndarray] = {}


def unicode_class_bitmaps() -> Tuple[np.ndarray, np.ndarray]:
    """(alnum, space) bitmaps, bit ``cp`` (little-endian within a byte) = ``chr(cp).isX()``."""
    if "alnum" not in _class_cache:
        alnum = np.zeros(UNICODE_LIMIT, dtype=np.uint8)
        space = np.zeros(UNICODE_LIMIT, dtype=np.uint8)
        for cp in range(UNICODE_LIMIT):
            ch = chr(cp)
            if ch.isalnum():
                alnum[cp] = 1
            elif ch.isspace():
                space[cp] = 1
        _class_cache["alnum"] = np.packbits(alnum, bitorder="little")
        _class_cache["space"] = np.packbits(space, bitorder="little")
    return _class_cache["alnum"], _class_cache["space"]


# str.isspace() code points the device pre-tokenizer hard-codes (csrc/pretok.cu); checked against the interpreter
PY_SPACE_CPS = frozenset([9, 10, 11, 12, 13, 28, 29, 30, 31, 32, 0x85, 0xA0, 0x1680, *range(0x2000, 0x200B), 0x2028, 0x2029,
                          0x202F, 0x205F, 0x3000])
LOWER_MULTI, LOWER_SIGMA, LOWER_PUNCT = 0x80000000, 0x40000000, 0x20000000
# whitespace of the Rust BertPreTokenizer (char::is_whitespace); the device kernel hard-codes it
RUST_SPACE_CPS = frozenset(PY_SPACE_CPS - {28, 29, 30, 31})
_bert_classes = None


def load_bert_classes() -> dict:
    """{'space': [[lo, hi], ...], 'punct': [[lo, hi], ...]} of the Rust BertPreTokenizer."""
    global _bert_classes
    if _bert_classes is None:
        import json
        import os
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "bert_pretok_classes.json")) as f:
            _bert_classes = json.load(f)
        spaces = {c for a, b in _bert_classes["space"] for c in range(a, b + 1)}
        if spaces != set(RUST_SPACE_CPS):
            raise RuntimeError("bert_pretok_classes.json: whitespace set differs from the one compiled into the device pre-tokenizer")
    return _bert_classes


def bert_pretokenizer_matches(pre) -> bool:
    """Spot check of the shipped class table against a live pre-tokenizer object (every code point below U+3100 and both
    ends of every range): False sends the caller to the host pre-tokenizer."""
    cls = load_bert_classes()
    kind = {}
    for name, code in (("space", 2), ("punct", 3)):
        for a, b in cls[name]:
            for c in range(a, b + 1):
                kind[c] = code
    probe = set(range(0x3100)) | {c for name in ("space", "punct") for a, b in cls[name] for c in (a - 1, a, b, b + 1)}
    for c in sorted(probe):
        if c < 0 or 0xD800 <= c < 0xE000 or c >= UNICODE_LIMIT:
            continue
        if len(pre.pre_tokenize_str("a" + chr(c) + "b")) != kind.get(c, 1):
            return False
    return True


class PretokTables:
    """Tables of the device pre-tokenizer of FastWP (``text.lower().split()``), generated from the RUNNING interpreter so
    that the device reproduces exactly this CPython's ``str.lower`` / ``str.isspace``.

    lower_map[cp]: lower-case code point | LOWER_MULTI|index into ``multi`` (one-to-many) | LOWER_SIGMA (U+03A3).
    ``sigma_bitmaps()``: Cased / Case_Ignorable bitmaps for the final-sigma rule, derived from ``str.lower`` itself
    (CPython does not expose the two properties): only needed when a text contains U+03A3.
    """

    _instance = None

    @classmethod
    def get(cls) -> "PretokTables":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def __init__(self):
        _, space = unicode_class_bitmaps()
        spaces = set(np.flatnonzero(np.unpackbits(space, bitorder="little")).tolist())
        if spaces != set(PY_SPACE_CPS):
            raise RuntimeError("str.isspace() of this interpreter differs from the set compiled into the device pre-tokenizer")
        lower = np.arange(UNICODE_LIMIT, dtype=np.uint32)
        multi: list = []
        last = 0
        for cp in range(0x80, UNICODE_LIMIT):
            if 0xD800 <= cp < 0xE000:
                continue
            lo = chr(cp).lower()
            if len(lo) == 1:
                if ord(lo) != cp:
                    lower[cp] = ord(lo); last = cp
            else:
                lower[cp] = LOWER_MULTI | len(multi); last = cp
                multi.append(len(lo)); multi.extend(ord(c) for c in lo)
                if len(lo) > 3:
                    raise RuntimeError("lower() of U+%04X has more than 3 code points" % cp)
        lower[0x3A3] = LOWER_SIGMA
        self._lower_full = lower
        self.lower_map = np.ascontiguousarray(lower[:max(last, 0x3A3) + 1])
        self.multi = np.asarray(multi, dtype=np.uint32)
        self._sigma = None
        self._bert = None

    def bert_lower_map(self) -> np.ndarray:
        """lower_map with bit 29 (LOWER_PUNCT) on the characters the Rust BertPreTokenizer isolates (data/bert_pretok_classes.json,
        probed from the `tokenizers` library by data/make_bert_classes.py)."""
        if self._bert is None:
            cls = load_bert_classes()
            m = self._lower_full.copy()
            top = len(self.lower_map)
            for a, b in cls["punct"]:
                m[a:b + 1] |= LOWER_PUNCT
                top = max(top, b + 1)
            self._bert = np.ascontiguousarray(m[:top])
        return self._bert

    def sigma_bitmaps(self) -> Tuple[np.ndarray, np.ndarray]:
        """(cased, case_ignorable) bitmaps.  Cased = islower|isupper|istitle of the single character.  Case_Ignorable is
        probed through the final-sigma rule of str.lower itself: for an uncased x, "a\u03a3" + x + "a" keeps a medial
        sigma iff x is ignorable; for a cased x, "1" + x + "\u03a3" yields a medial sigma iff x is ignorable."""
        if self._sigma is None:
            cased = np.zeros(UNICODE_LIMIT, dtype=np.uint8)
            ign = np.zeros(UNICODE_LIMIT, dtype=np.uint8)
            for cp in range(UNICODE_LIMIT):
                if 0xD800 <= cp < 0xE000:
                    continue
                ch = chr(cp)
                if ch.islower() or ch.isupper() or ch.istitle():
                    cased[cp] = 1
                    if ("1" + ch + "\u03a3").lower()[-1] == "\u03c3":
                        ign[cp] = 1
                elif ("a\u03a3" + ch + "a").lower()[1] == "\u03c3":
                    ign[cp] = 1
            self._sigma = (np.packbits(cased, bitorder="little"), np.packbits(ign, bitorder="little"))
        return self._sigma


def encode_utf8(word: str) -> bytes:
    # lone surrogates are legal in a Python str; keep them round-trippable
    return word.encode("utf-8", "surrogatepass")


def decode_utf8(b: bytes) -> str:
    return b.decode("utf-8", "surrogatepass")


def pack_words(words: Sequence[str], offset_dtype=np.uint64) -> Tuple[np.ndarray, np.ndarray]:
    """Byte arena + ``len(words)+1`` offsets."""
    enc = [encode_utf8(w) for w in words]
    lens = np.fromiter((len(b) for b in enc), dtype=np.uint64, count=len(enc))
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    arena = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if enc else np.zeros(0, dtype=np.uint8)
    return arena, off.astype(offset_dtype, copy=False)


def cps_of(s: str) -> np.ndarray:
    return np.frombuffer(s.encode("utf-32-le", "surrogatepass"), dtype=np.uint32)


def pack_strings_as_cps(strings: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Code-point arena (u32) + offsets (u64) for a list of strings."""
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    if strings:
        np.cumsum(np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings)), out=off[1:])
    cps = np.frombuffer("".join(strings).encode("utf-32-le", "surrogatepass"), dtype=np.uint32).copy()
    return cps, off


class BpeTables:
    """Integer form of ``merges_list`` / ``_bpe_ranks`` (reference bpe.py:196,200,257).

    ``id_to_str[i]`` is the string of symbol ``i``.  Ids are canonical per distinct string, in
    order of first appearance while scanning merge ``k`` as (left, right, left+right).
    """

    def __init__(self, merges: Sequence[Tuple[str, str]]):
        str_to_id: Dict[str, int] = {}
        id_to_str: List[str] = []

        def intern(s: str) -> int:
            i = str_to_id.get(s)
            if i is None:
                i = len(id_to_str)
                str_to_id[s] = i
                id_to_str.append(s)
            return i

        n = len(merges)
        left = np.zeros(n, dtype=np.uint32)
        right = np.zeros(n, dtype=np.uint32)
        new = np.zeros(n, dtype=np.uint32)
        for k, pair in enumerate(merges):
            a, b = pair[0], pair[1]
            left[k] = intern(a)
            right[k] = intern(b)
            new[k] = intern(a + b)
        self.left, self.right, self.new = left, right, new
        self.id_to_str = id_to_str
        self.str_to_id = str_to_id
        chars = sorted((ord(s), i) for s, i in str_to_id.items() if len(s) == 1)
        self.char_cp = np.array([c for c, _ in chars], dtype=np.uint32)
        self.char_id = np.array([i for _, i in chars], dtype=np.uint32)

    @property
    def n_merges(self) -> int:
        return int(self.left.shape[0])

    def token_to_str(self, tok: int) -> str:
        """Inverse of the C-ABI token encoding ``(symbol << 1) | continuation``."""
        if tok == BPE_EMPTY_TOKEN:
            return ""
        sym, cont = tok >> 1, tok & 1
        s = chr(sym & 0x1FFFFF) if sym & BPE_UNKNOWN_CP else self.id_to_str[sym]
        return "##" + s if cont else s

    def tokens_to_strs(self, toks: Iterable[int]) -> List[str]:
        """Token ids -> the reference's token strings.  Vectorised: one fancy-indexing pass over an object array of
        [symbol, "##" + symbol] pairs; ids outside it (characters no merge mentions, the empty word) take the scalar path."""
        toks = np.asarray(toks, dtype=np.uint32) if not isinstance(toks, np.ndarray) else toks.astype(np.uint32, copy=False)
        if toks.size == 0:
            return []
        lut = getattr(self, "_str_lut", None)
        if lut is None:
            lut = np.empty(2 * len(self.id_to_str), dtype=object)
            lut[0::2] = self.id_to_str
            lut[1::2] = ["##" + x for x in self.id_to_str]
            self._str_lut = lut
        try:                                   # the common case in one call (per-line tokenize() is a few microseconds of Python in total)
            return lut[toks].tolist()
        except IndexError:                     # ids outside the table: characters no merge mentions (0x40000000 | cp), the empty word
            pass
        known = toks < len(lut)
        if known.all():
            return lut[toks].tolist()
        out = np.empty(toks.size, dtype=object)
        out[known] = lut[toks[known]]
        for k in np.flatnonzero(~known):
            out[k] = self.token_to_str(int(toks[k]))
        return out.tolist()


class WpTables:
    """Sorted vocabulary as code points; token id = index, ``n`` = "['UNK']", ``n+1`` = "[UNK]"."""

    UNK_FAST = "['UNK']"   # reference wordpiece.py:257
    UNK_NAIVE = "[UNK]"    # reference wordpiece.py:149

    def __init__(self, vocab: Iterable[str]):
        self.tokens: List[str] = sorted(set(vocab))
        for t in self.tokens:
            if any(ch.isspace() for ch in t):
                # reference wordpiece.py:285 indexes past the end / crosses words for such vocabularies
                raise NotImplementedError("vocabulary entries containing whitespace are outside the FastWP parity domain")
        self.cps, self.off = pack_strings_as_cps(self.tokens)
        self.id_to_str: List[str] = self.tokens + [self.UNK_FAST, self.UNK_NAIVE]

    @property
    def n_vocab(self) -> int:
        return len(self.tokens)

    def tokens_to_strs(self, toks: Iterable[int]) -> List[str]:
        lut = getattr(self, "_str_lut", None)
        if lut is None:
            lut = self._str_lut = np.array(self.id_to_str, dtype=object)
        if not isinstance(toks, np.ndarray):
            toks = np.asarray(toks, dtype=np.int64)
        return lut[toks].tolist() if toks.size else []


class TrainTypes:
    """Word types of a pre-tokenized corpus in first-occurrence order (reference bpe.py:73-81)."""

    def __init__(self, words: Sequence[str]):
        freqs = Counter(words)                        # insertion order == first occurrence
        self.types: List[str] = list(freqs.keys())
        self.freq = np.fromiter(freqs.values(), dtype=np.int64, count=len(freqs))
        alphabet = sorted({ch for w in self.types for ch in w})
        self.alphabet: List[str] = alphabet
        cp_to_id = {ord(c): i for i, c in enumerate(alphabet)}
        cps, off = pack_strings_as_cps(self.types)
        if len(cps):
            lut_keys = np.fromiter(cp_to_id.keys(), dtype=np.uint32, count=len(cp_to_id))
            lut_vals = np.fromiter(cp_to_id.values(), dtype=np.uint32, count=len(cp_to_id))
            order = np.argsort(lut_keys)
            lut_keys, lut_vals = lut_keys[order], lut_vals[order]
            self.syms = lut_vals[np.searchsorted(lut_keys, cps)].astype(np.uint32)
        else:
            self.syms = np.zeros(0, dtype=np.uint32)
        self.off = off

    def __getattr__(self, name):
        # instances built on the device (device.device_train_types) materialise the type strings on demand
        if name == "types" and "_type_strings" in self.__dict__:
            self.types = self._type_strings()
            return self.types
        raise AttributeError(name)

    @property
    def n_types(self) -> int:
        return len(self.freq)

    @property
    def n_alpha(self) -> int:
        return len(self.alphabet)

    def merges_to_strs(self, left, right, new) -> Tuple[List[Tuple[str, str]], List[str]]:
        """(left,right,new) id triples -> (merges_list, symbol strings)."""
        strs: List[str] = list(self.alphabet)
        merges: List[Tuple[str, str]] = []
        for a, b, z in zip(left.tolist(), right.tolist(), new.tolist()):
            merges.append((strs[a], strs[b]))
            if z == len(strs):
                strs.append(strs[a] + strs[b])
            elif z > len(strs):
                raise ValueError("merged id %d skips ahead of the symbol table (%d)" % (z, len(strs)))
        return merges, strs


class WpTrainTypes:
    """Word types for NaiveWP.train (reference wordpiece.py:49-62): first char bare, the rest
    prefixed with ``##``; one id per distinct initial symbol string."""

    def __init__(self, words: Sequence[str]):
        freqs = Counter(words)
        self.types: List[str] = list(freqs.keys())
        self.freq = np.fromiter(freqs.values(), dtype=np.int64, count=len(freqs))
        sym_to_id: Dict[str, int] = {}
        syms: List[int] = []
        off = [0]
        for w in self.types:
            for k, c in enumerate(w):
                s = c if k == 0 else "##" + c
                i = sym_to_id.get(s)
                if i is None:
                    i = sym_to_id[s] = len(sym_to_id)
                syms.append(i)
            off.append(len(syms))
        self.init_syms: List[str] = list(sym_to_id.keys())
        self.syms = np.array(syms, dtype=np.uint32)
        self.off = np.array(off, dtype=np.uint64)
        self.init_cps, self.init_off = pack_strings_as_cps(self.init_syms)

    def __getattr__(self, name):
        if name == "types" and "_type_strings" in self.__dict__:
            self.types = self._type_strings()
            return self.types
        raise AttributeError(name)

    def vocab_from_merges(self, left, right, new) -> List[str]:
        strs: List[str] = list(self.init_syms)
        for a, b, z in zip(left.tolist(), right.tolist(), new.tolist()):
            if z == len(strs):
                strs.append(strs[a] + strs[b][2:])        # wordpiece.py:95
        return strs
