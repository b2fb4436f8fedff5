"""NaiveWP / FastWP with the reference's class surface (source/wordpiece.py).

* ``FastWP.tokenize`` -> HP-2 kernel (swt_wp_encode) over the device trie.  reference wordpiece.py:233-316
* ``NaiveWP`` (trainer with score freq/(f_a*f_b), greedy longest-prefix encoder) is not on a north-star
  hot path; it stays a host implementation so that FastWP.train and
  --compare keep working.
"""
from __future__ import annotations

import json
import os
from collections import Counter
from typing import Dict, List, Sequence, Tuple

from .utils import SubwordTokenizer, WPTrie_E2E, naive_wp_encode


class NaiveWP(SubwordTokenizer):
    """WordPiece tokenizer (reference source/wordpiece.py:8-208)."""

    def __init__(self, tokenizer):
        super().__init__(tokenizer)
        self.vocab: set = set()
        self.corpus_as_symbols: List[Tuple[List[str], int]] = []

    def train(self, corpus, max_vocab: int = 30_000):
        if not isinstance(corpus, list) or not all(isinstance(example, str) for example in corpus):
            raise TypeError("corpus must be a list of strings.")
        if not isinstance(max_vocab, int):
            raise TypeError("max_vocab must be an int.")
        self.reset()
        word_freqs = Counter(self._pre_tokenized_words(corpus))
        words = [([w[0]] + ["##" + c for c in w[1:]], f) for w, f in word_freqs.items()]
        self.corpus_as_symbols.extend(words)
        self.vocab |= {s for symbols, _ in words for s in symbols}
        # Host loop with incrementally maintained counts; selection rule of wordpiece.py:84-92:
        # score = pair_freq / (freq_a * freq_b) as a Python float, first-inserted pair wins ties.
        while len(self.vocab) < max_vocab:
            pair_freqs: Dict[Tuple[str, str], int] = {}
            sym_freqs: Dict[str, int] = {}
            for symbols, f in self.corpus_as_symbols:
                prev = None
                for s in symbols:
                    sym_freqs[s] = sym_freqs.get(s, 0) + f
                    if prev is not None:
                        key = (prev, s)
                        pair_freqs[key] = pair_freqs.get(key, 0) + f
                    prev = s
            if not pair_freqs:
                break
            best, best_score = None, -1.0
            for pair, f in pair_freqs.items():
                score = f / (sym_freqs[pair[0]] * sym_freqs[pair[1]])
                if score > best_score:
                    best, best_score = pair, score
            self.vocab.add(best[0] + best[1][2:])
            self.corpus_as_symbols = [(self._replace_pair(best, symbols), f) for symbols, f in self.corpus_as_symbols]

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

    def tokenize(self, text):
        if not isinstance(text, str):
            raise TypeError("Text to tokenize must be a string.")
        out: List[str] = []
        for word in self._pre_tokenized_words([text]):
            out.extend(self.encode_word(word))
        return out

    def reset(self) -> None:
        self.vocab.clear()
        self.corpus_as_symbols.clear()

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

    @staticmethod
    def _chunks(text: str) -> List[str]:
        # s = text.lower() + " " (wordpiece.py:248); whitespace (Python str.isspace, :268) only ever separates
        # segments, so the device works on the whitespace-free chunks.  str.split() splits on exactly the
        # characters for which str.isspace() is true.
        return text.lower().split()

    def tokenize(self, text):
        if not isinstance(text, str):
            raise TypeError("Text to tokenize must be a string.")
        trie = self.vocab_trie                      # AttributeError before train/load, like the reference
        ids, _, _ = trie.encoder.encode_words(self._chunks(text))
        return trie.tables.tokens_to_strs(ids)

    def tokenize_batch(self, texts: Sequence[str]) -> List[List[str]]:
        trie = self.vocab_trie
        per_text = [self._chunks(t) for t in texts]
        ids, tok_off, _ = trie.encoder.encode_words([c for cs in per_text for c in cs])
        strs = trie.tables.tokens_to_strs(ids)
        out, wi = [], 0
        for cs in per_text:
            out.append(strs[int(tok_off[wi]):int(tok_off[wi + len(cs)])])
            wi += len(cs)
        return out

    def load_resources(self, path: str) -> None:
        super().load_resources(path)
        self.vocab_trie = WPTrie_E2E(self.vocab)

    def save_resources(self, path: str) -> None:
        super().save_resources(path)
