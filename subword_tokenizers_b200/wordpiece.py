"""NaiveWP / FastWP with the reference's class surface (source/wordpiece.py).

* ``FastWP.tokenize`` -> HP-2 kernel (swt_wp_encode) over the device trie.  reference wordpiece.py:233-316
* ``NaiveWP.train`` (score freq/(f_a*f_b), reference wordpiece.py:29-103) is the first "next" row of SURVEY.md §8(f):
  it runs on the GPU through the WordPiece mode of the trainer kernels.  ``NaiveWP.encode_word`` (greedy
  longest-prefix, wordpiece.py:131-158) stays a host loop; it is not a hot path and exists for --compare.
"""
from __future__ import annotations

import json
import os
from collections import Counter
from typing import Dict, List, Sequence, Tuple

from . import packing as P
from ._tracked import TrackedSet, stamp, tracked_attribute
from .utils import SubwordTokenizer, WPTrie_E2E, naive_wp_encode


class NaiveWP(SubwordTokenizer):
    """WordPiece tokenizer (reference source/wordpiece.py:8-208)."""

    vocab = tracked_attribute("vocab", TrackedSet)      # counts its mutations: the device trie follows the live set

    def __init__(self, tokenizer):
        super().__init__(tokenizer)
        self.vocab: set = set()
        self._corpus_cache: List[Tuple[List[str], int]] = []
        self._train_result = None

    def train(self, corpus, max_vocab: int = 30_000):
        if not isinstance(corpus, list) or not all(isinstance(example, str) for example in corpus):
            raise TypeError("corpus must be a list of strings.")
        if not isinstance(max_vocab, int):
            raise TypeError("max_vocab must be an int.")
        self.reset()
        if self._device_pretok_ok():
            from .device import device_train_types
            self._train_on_types(device_train_types(corpus, wordpiece=True), max_vocab)   # pre-tokenization + counting on the GPU
        else:
            self.train_on_words(self._pre_tokenized_words(corpus), max_vocab)

    def train_on_words(self, words: Sequence[str], max_vocab: int) -> None:
        from . import packing as P
        self._train_on_types(P.WpTrainTypes(words), max_vocab)

    def _train_on_types(self, types, max_vocab: int) -> None:
        """The merge loop of wordpiece.py:68-102 on the GPU (SWT_TRAIN_WP mode of the trainer): score
        pair_freq / (freq_a * freq_b), first-inserted pair on ties, merged token a + b[2:]."""
        import numpy as np
        import torch
        from . import _lib, packing as P
        from .device import CudaTrainEngine, run_training_loop
        self.vocab = set(types.init_syms)
        if len(types.freq) == 0 or len(types.init_syms) == 0:      # empty corpus: nothing to merge (wordpiece.py:68,74-75)
            self._train_result = None
            self._corpus_cache = []
            return
        max_len = int(np.diff(types.off.astype(np.int64)).max()) if len(types.freq) else 1
        engine = CudaTrainEngine(types.syms, types.off, types.freq, len(types.init_syms), max_vocab, len(types.init_syms),
                                 max_len + 2, 0, 0, 1, mode=_lib.TRAIN_WP, init_cps=types.init_cps, init_off=types.init_off)
        left, right, new, count, state = run_training_loop(engine, 1)
        strs = types.vocab_from_merges(left, right, new)
        self.vocab = set(strs)
        self._train_result = (engine, types, strs)
        self._corpus_cache = None

    @property
    def corpus_as_symbols(self) -> List[Tuple[List[str], int]]:
        """(symbols, freq) per word type (reference wordpiece.py:27,99-102), read back lazily from the device."""
        if self._corpus_cache is None:
            engine, types, strs = self._train_result
            syms, lens = engine.read_corpus()
            out = []
            for k in range(len(types.freq)):
                s0 = int(types.off[k])
                out.append(([strs[i] for i in syms[s0:s0 + int(lens[k])]], int(types.freq[k])))
            self._corpus_cache = out
        return self._corpus_cache

    @corpus_as_symbols.setter
    def corpus_as_symbols(self, value) -> None:
        self._corpus_cache = value

    def _replace_pair(self, pair, word):
        merged = pair[0] + pair[1][2:]
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

    def encode_word(self, word):
        return naive_wp_encode(word, self.vocab)

    def _naive_device_encoder(self):
        """Greedy longest-prefix encoder on the device (swt_wp_encode_naive) over the trie of the current vocabulary."""
        from .device import WpEncoder
        from .utils import naive_wp_encode_ids
        quick = stamp(self.vocab)
        if getattr(self, "_naive_quick", None) != quick:
            tables = P.WpTables(self.vocab)
            self._naive_encoder = WpEncoder(tables, naive_wp_encode_ids("##", tables), naive=True)
            self._naive_quick = quick
        return self._naive_encoder

    def encode_words(self, words: Sequence[str]) -> List[List[str]]:
        """Batch form of encode_word: one kernel launch for all words."""
        enc = self._naive_device_encoder()
        ids, tok_off, _ = enc.encode_words(words)
        strs = enc.tables.tokens_to_strs(ids)
        return [strs[int(tok_off[i]):int(tok_off[i + 1])] for i in range(len(words))]

    def tokenize_batch(self, texts: Sequence[str]) -> List[List[str]]:
        """tokenize() of every text with one pass over their concatenation."""
        from .bpe import _batch_lists
        return _batch_lists(self, self._naive_device_encoder(), texts)

    def tokenize(self, text):
        if not isinstance(text, str):
            raise TypeError("Text to tokenize must be a string.")
        enc = self._naive_device_encoder()
        if self._device_pretok_ok():
            ids = enc.encode_text(text)                 # lower-casing + BERT pre-tokenization + greedy matching on the device
        else:
            ids, _, _ = enc.encode_words(self._pre_tokenized_words([text]))
        return enc.tables.tokens_to_strs(ids)

    def reset(self) -> None:
        self.vocab.clear()
        self._corpus_cache = []
        self._train_result = None

    def save_resources(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "vocab.json"), "w", encoding="utf-8") as f:
            json.dump(list(self.vocab), f, ensure_ascii=False)

    def load_resources(self, path: str) -> None:
        vocab_file = os.path.join(path, "vocab.json")
        if os.path.isfile(vocab_file):
            with open(vocab_file, "r", encoding="utf-8") as f:
                self.vocab = set(json.load(f))


class FastWP(NaiveWP):
    """End-to-end LinMaxMatch WordPiece (reference source/wordpiece.py:211-330) on the GPU."""

    def __init__(self, tokenizer):
        super().__init__(tokenizer)

    def train(self, corpus, max_vocab=30_000):
        super().train(corpus, max_vocab)
        self.vocab_trie = WPTrie_E2E(self.vocab)

    def tokenize(self, text):
        if not isinstance(text, str):
            raise TypeError("Text to tokenize must be a string.")
        trie = self.vocab_trie                      # AttributeError before train/load, like the reference
        ids = trie.encoder.encode_text(text)        # lower-casing + whitespace split + LinMaxMatch, all on the device
        return trie.tables.tokens_to_strs(ids)

    def tokenize_batch(self, texts: Sequence[str]) -> List[List[str]]:
        """tokenize() of every text with one pass over their concatenation (lower-casing, splitting and matching on the device)."""
        if not all(isinstance(t, str) for t in texts):
            raise TypeError("Text to tokenize must be a string.")
        trie = self.vocab_trie
        ids, cut = trie.encoder.encode_texts(texts)
        strs = trie.tables.tokens_to_strs(ids)
        return [strs[int(cut[k]):int(cut[k + 1])] for k in range(len(texts))]

    def load_resources(self, path: str) -> None:
        super().load_resources(path)
        self.vocab_trie = WPTrie_E2E(self.vocab)

    def save_resources(self, path: str) -> None:
        super().save_resources(path)
