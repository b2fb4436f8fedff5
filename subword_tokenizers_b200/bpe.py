"""NaiveBPE / FastBPE with the reference's class surface (source/bpe.py), computing on the GPU.

* ``train``       -> HP-3 kernels (swt_bpe_train_*): pair counts, argmax with the reference's
                     first-inserted tie-break, in-place merge-apply.     reference bpe.py:50-112
* ``FastBPE.encode_word`` / ``tokenize`` -> HP-1 kernel (swt_bpe_encode).  reference bpe.py:205-249
* ``NaiveBPE.encode_word`` (in-order merge replay, bpe.py:114-132) stays a host loop: it is not a hot
  path and exists for --compare.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import packing as P
from ._tracked import TrackedDict, TrackedList, TrackedSet, stamp, tracked_attribute
from .utils import SubwordTokenizer


def _batch_lists(tok, enc, texts: Sequence[str]) -> List[List[str]]:
    """Shared body of tokenize_batch: device batch call when the pre-tokenizer is the BERT one, else host pre-tokenization."""
    if not all(isinstance(t, str) for t in texts):
        raise TypeError("Text must be a string.")
    if tok._device_pretok_ok():
        ids, cut = enc.encode_texts(texts)
    else:
        pre = tok.tokenizer.backend_tokenizer.pre_tokenizer
        per_text = [[w for w, _ in pre.pre_tokenize_str(t.lower())] for t in texts]
        ids, tok_off, _ = enc.encode_words([w for ws in per_text for w in ws])
        cut = tok_off.astype(np.int64)[np.cumsum([0] + [len(ws) for ws in per_text])]
    strs = enc.tables.tokens_to_strs(ids)
    return [strs[int(cut[k]):int(cut[k + 1])] for k in range(len(texts))]


class NaiveBPE(SubwordTokenizer):
    """Byte-Pair-Encoding tokenizer (reference source/bpe.py:9-189)."""

    # mutation-counting containers: the device tables are rebuilt whenever the live Python object changed (_tracked.py)
    merges_list = tracked_attribute("merges_list", TrackedList)
    vocab = tracked_attribute("vocab", TrackedSet)

    def __init__(self, tokenizer) -> None:
        super().__init__(tokenizer)
        self.merges_list: List[Tuple[str, str]] = []
        self.vocab: set = set()
        self._corpus_cache: Optional[List[Tuple[List[str], int]]] = []
        self._train_result = None
        self.last_train_stats: Dict[str, float] = {}

    # -- training (HP-3) -------------------------------------------------------------------------------------
    def train(self, corpus: List[str], max_vocab: int = 30_000) -> None:
        if not isinstance(corpus, list) or not all(isinstance(example, str) for example in corpus):
            raise TypeError("Corpus must be a list of strings.")
        if not isinstance(max_vocab, int):
            raise TypeError("Maximum vocabulary size must be an integer.")
        self.reset()
        if self._device_pretok_ok():
            from .device import device_train_types
            self._train_on_types(device_train_types(corpus, wordpiece=False), max_vocab)   # pre-tokenization + counting on the GPU
        else:
            self.train_on_words(self._pre_tokenized_words(corpus), max_vocab)

    def train_on_words(self, words: Sequence[str], max_vocab: int, group=None) -> None:
        self._train_on_types(P.TrainTypes(words), max_vocab, group)

    def _train_on_types(self, types, max_vocab: int, group=None) -> None:
        """The merge loop on pre-tokenized words.  With torch.distributed initialised (world > 1) the
        word types are sharded across ranks and every rank returns the same merge list."""
        import time
        import torch
        import torch.distributed as dist
        from .device import CudaTrainEngine, run_training_loop, shard_types

        self.vocab = set(types.alphabet)
        if types.n_types == 0 or types.n_alpha == 0:          # empty corpus: the reference returns with no merges (bpe.py:88,98-99)
            self.merges_list = []
            self._train_result = None
            self._corpus_cache = []
            self.last_train_stats = {"merge_loop_s": 0.0, "merges": 0, "n_types": 0, "n_symbols": 0, "world_size": 1}
            return
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        t0, t1 = shard_types(types.off, world)[rank]
        off = types.off[t0:t1 + 1] - types.off[t0]
        syms = types.syms[int(types.off[t0]):int(types.off[t1])]
        max_len = int(np.diff(types.off.astype(np.int64)).max()) if types.n_types else 1
        # alphabets beyond the dense-count limit insert their initial pairs directly: a sharded table must then be sized
        # for the pairs of ALL ranks from the start
        table_cap = 2 * int(len(types.syms)) + 8 * max(max_vocab, 1) if (world > 1 and types.n_alpha > 4096) else 0
        engine = CudaTrainEngine(syms, off, types.freq[t0:t1], types.n_alpha, max_vocab, len(types.alphabet), max_len,
                                 int(types.off[t0]), rank, world, table_cap=table_cap)
        torch.cuda.synchronize()
        tic = time.perf_counter()
        left, right, new, count, state = run_training_loop(engine, world, group)
        torch.cuda.synchronize()
        self.last_train_stats = {"merge_loop_s": time.perf_counter() - tic, "merges": int(len(left)),
                                 "n_types": types.n_types, "n_symbols": int(len(types.syms)), "world_size": world}
        merges, strs = types.merges_to_strs(left, right, new)
        self.merges_list = merges
        self.vocab.update(strs[types.n_alpha:])
        self._train_result = (engine, types, strs, t0, t1)
        self._corpus_cache = None

    @property
    def corpus_as_symbols(self) -> List[Tuple[List[str], int]]:
        """(symbols, freq) per word type after training (reference bpe.py:23,108-111); materialised
        lazily from the device word table (this rank's shard when training was sharded)."""
        if self._corpus_cache is None:
            engine, types, strs, t0, t1 = self._train_result
            syms, lens = engine.read_corpus()
            out = []
            starts = (types.off[t0:t1] - types.off[t0]).astype(np.int64)
            for k in range(t1 - t0):
                s = int(starts[k])
                out.append(([strs[i] for i in syms[s:s + int(lens[k])]], int(types.freq[t0 + k])))
            self._corpus_cache = out
        return self._corpus_cache

    @corpus_as_symbols.setter
    def corpus_as_symbols(self, value) -> None:
        self._corpus_cache = value

    # -- encoding ----------------------------------------------------------------------------------------------
    def _replace_pair(self, pair: Tuple[str, str], word: List[str]) -> List[str]:
        merged = pair[0] + pair[1]
        out: List[str] = []
        k, n = 0, len(word)
        while k < n:
            if k + 1 < n and word[k] == pair[0] and word[k + 1] == pair[1]:
                out.append(merged)
                k += 2
            else:
                out.append(word[k])
                k += 1
        return out

    def encode_word(self, word: str) -> List[str]:
        pieces = list(word)
        for pair in self.merges_list:
            pieces = self._replace_pair(pair, pieces)
        return pieces[:1] + ["##" + p for p in pieces[1:]]

    def _naive_device_encoder(self):
        """In-order replay on the device (swt_bpe_encode_naive), or None when the merge list repeats a pair (the device table
        keeps one rank per pair; the host replay below stays exact for such lists)."""
        from .device import BpeEncoder
        quick = stamp(self.merges_list)
        if getattr(self, "_naive_quick", None) != quick:
            pairs = [tuple(p) for p in self.merges_list]
            self._naive_encoder = BpeEncoder(P.BpeTables(pairs), naive=True) if len(set(pairs)) == len(pairs) else None
            self._naive_quick = quick
        return self._naive_encoder

    def encode_words(self, words: Sequence[str]) -> List[List[str]]:
        """Batch form of encode_word: one kernel launch for all words (host replay when the merge list repeats a pair)."""
        enc = self._naive_device_encoder()
        if enc is None:
            return [self.encode_word(w) for w in words]
        ids, tok_off, _ = enc.encode_words(words)
        strs = enc.tables.tokens_to_strs(ids)
        return [strs[int(tok_off[i]):int(tok_off[i + 1])] for i in range(len(words))]

    def tokenize_batch(self, texts: Sequence[str]) -> List[List[str]]:
        """tokenize() of every text with one pass over their concatenation."""
        enc = self._naive_device_encoder()
        if enc is None:
            return [self.tokenize(t) for t in texts]
        return _batch_lists(self, enc, texts)

    def tokenize(self, text: str) -> List[str]:
        if not isinstance(text, str):
            raise TypeError("Text to tokenize must be a string.")
        enc = self._naive_device_encoder()
        if enc is not None:
            if self._device_pretok_ok():
                ids = enc.encode_text(text)
            else:
                ids, _, _ = enc.encode_words(self._pre_tokenized_words([text]))
            return enc.tables.tokens_to_strs(ids)
        out: List[str] = []
        for word in self._pre_tokenized_words([text]):
            out.extend(self.encode_word(word))
        return out

    def reset(self) -> None:
        self.merges_list.clear()
        self.vocab.clear()
        self._corpus_cache = []
        self._train_result = None

    # -- persistence (on-disk layout of reference bpe.py:167-189) ------------------------------------------
    def save_resources(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "merges.json"), "w", encoding="utf-8") as f:
            json.dump(self.merges_list, f, ensure_ascii=False)

    def load_resources(self, path: str) -> None:
        merges_file = os.path.join(path, "merges.json")
        if os.path.isfile(merges_file):                # a missing file is silently ignored, like the reference
            with open(merges_file, "r", encoding="utf-8") as f:
                self.merges_list = [tuple(pair) for pair in json.load(f)]


class FastBPE(NaiveBPE):
    """Rank-map BPE inference (reference source/bpe.py:192-263) on the GPU."""

    _bpe_ranks = tracked_attribute("_bpe_ranks", TrackedDict)

    def __init__(self, tokenizer):
        super().__init__(tokenizer)
        self._bpe_ranks = {}
        self._encoder = None
        self._encoder_quick = None

    def _rebuild_ranks(self) -> None:
        self._bpe_ranks = {pair: i for i, pair in enumerate(self.merges_list)}
        self._encoder = None

    def train(self, corpus: List[str], max_vocab: int = 30_000) -> None:
        super().train(corpus, max_vocab)
        self._rebuild_ranks()

    def train_on_words(self, words, max_vocab, group=None) -> None:
        super().train_on_words(words, max_vocab, group)
        self._rebuild_ranks()

    def _device_encoder(self):
        """The rank table on the device, rebuilt when merges_list / _bpe_ranks changed."""
        from .device import BpeEncoder
        # honour direct assignment to / in-place edits of _bpe_ranks (the reference reads only that dict in encode_word):
        # the dict counts its mutations, so the check is O(1) on the per-call path and still exact
        quick = stamp(self._bpe_ranks)
        if self._encoder is not None and self._encoder_quick == quick:
            return self._encoder
        ranked = sorted(self._bpe_ranks.items(), key=lambda kv: kv[1])
        self._encoder = BpeEncoder(P.BpeTables([p for p, _ in ranked]))
        self._encoder_quick = quick
        return self._encoder

    def _pairs(self, seq: List[str]) -> set:
        return {(seq[i], seq[i + 1]) for i in range(len(seq) - 1)}

    def encode_words(self, words: Sequence[str]) -> List[List[str]]:
        """Batch form of encode_word: one kernel launch for all words."""
        enc = self._device_encoder()
        ids, tok_off, _ = enc.encode_words(words)
        strs = enc.tables.tokens_to_strs(ids)
        return [strs[int(tok_off[i]):int(tok_off[i + 1])] for i in range(len(words))]

    def encode_word(self, word: str) -> List[str]:
        return self.encode_words([word])[0]

    def tokenize(self, text: str) -> List[str]:
        if not isinstance(text, str):
            raise TypeError("Text must be a string.")
        enc = self._device_encoder()
        if self._device_pretok_ok():
            ids = enc.encode_text(text)             # lower-casing + BERT pre-tokenization + merge loop, all on the device
        else:
            ids, _, _ = enc.encode_words(self._pre_tokenized_words([text]))
        return enc.tables.tokens_to_strs(ids)

    def tokenize_batch(self, texts: Sequence[str]) -> List[List[str]]:
        """tokenize() of every text with one pass over their concatenation (pre-tokenization and merge loop on the device);
        returns one token list per text."""
        enc = self._device_encoder()
        return _batch_lists(self, enc, texts)

    def load_resources(self, path: str) -> None:
        super().load_resources(path)
        self._rebuild_ranks()

    def save_resources(self, path: str) -> None:
        super().save_resources(path)
