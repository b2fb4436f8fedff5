"""Base class and trie facade mirroring the reference's source/utils.py surface."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple


class SubwordTokenizer:
    """Parent of the four tokenizers (reference source/utils.py:5-41).

    ``tokenizer`` is any object exposing ``backend_tokenizer.pre_tokenizer.pre_tokenize_str``;
    only the BERT pre-tokenizer is used, after Python lower-casing (utils.py:26-29).
    """

    def __init__(self, tokenizer) -> None:
        self.tokenizer = tokenizer

    def preprocessing(self, corpus: List[str]) -> List[List[Tuple[str, Tuple[int, int]]]]:
        pre = self.tokenizer.backend_tokenizer.pre_tokenizer
        return [pre.pre_tokenize_str(example.lower()) for example in corpus]

    def _pre_tokenized_words(self, corpus: List[str]) -> List[str]:
        pre = self.tokenizer.backend_tokenizer.pre_tokenizer
        return [w for example in corpus for w, _ in pre.pre_tokenize_str(example.lower())]

    def _device_pretok_ok(self) -> bool:
        """The device pre-tokenizer reproduces the Rust BertPreTokenizer (the only one the reference ever uses, cli.py:163);
        any other pre-tokenizer object, or a `tokenizers` build whose character classes differ from the shipped table, keeps
        the host call of utils.py:27."""
        ok = getattr(self, "_pretok_checked", None)
        if ok is None:
            from . import packing as P
            pre = self.tokenizer.backend_tokenizer.pre_tokenizer
            ok = type(pre).__name__ == "BertPreTokenizer" and P.bert_pretokenizer_matches(pre)
            self._pretok_checked = ok
        return ok

    def vocab_length(self, corpus: List[str]) -> int:
        return len({symbol for example in corpus for symbol in example})


class WPTrie_E2E:
    """Handle of the device-resident end-to-end WordPiece trie (reference source/utils.py:66-139).

    The reference builds Python TrieNode objects; here insertion, the failure-link / failure-pop
    precompute and the punctuation rule run in libswt (swt_wp_trie_create) and the result lives in HBM.
    ``stats()`` reports nodes / edges / pops / links to root_p for inspection.
    """

    def __init__(self, vocab: Iterable[str] = ()):
        from . import packing as P
        from .device import WpEncoder
        self.tables = P.WpTables(vocab)
        self.sharp_special = naive_wp_encode_ids("##", self.tables)
        self.encoder = WpEncoder(self.tables, self.sharp_special)

    def stats(self):
        return self.encoder.stats()


def naive_wp_encode(word: str, vocab) -> List[str]:
    """Greedy longest-prefix WordPiece of one word (behaviour of reference wordpiece.py:131-158).

    When "#" is a vocabulary entry but the "##"-prefixed remainder has no match, the reference's
    remainder grows without bound; that case is reported as ["[UNK]"] (parity domain).
    """
    tokens: List[str] = []
    rest = word
    while rest:
        end = len(rest)
        while end > 0 and rest[:end] not in vocab:
            end -= 1
        if end == 0:
            return ["[UNK]"]
        tokens.append(rest[:end])
        tail = rest[end:]
        if not tail:
            break
        if len(tail) + 2 >= len(rest) and len(tokens) > 1:
            return ["[UNK]"]
        rest = "##" + tail
    return tokens


def naive_wp_encode_ids(word: str, tables) -> List[int]:
    index = {t: i for i, t in enumerate(tables.id_to_str)}
    return [index[t] for t in naive_wp_encode(word, set(tables.tokens))]
