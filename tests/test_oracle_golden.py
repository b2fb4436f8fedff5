"""Pins the CPU oracle (oracle/swt_oracle.c) to the reference: golden vectors shipped with the
reference (data/pan_tadeusz.tokens.json, resources/tests KATs), the survey-time hashes, and
randomized cases whose expected outputs were produced by running the unmodified Python reference
(tests/golden/make_golden.py).  No GPU needed."""
import hashlib
import json

import numpy as np
import pytest

import oracle
from conftest import load_golden
from subword_tokenizers_b200 import packing as P


def _split_by_offsets(strs, tok_off, counts):
    out, wi = [], 0
    for n in counts:
        out.append(strs[int(tok_off[wi]):int(tok_off[wi + n])])
        wi += n
    return out


def _bpe_train_oracle(words, max_vocab):
    tt = P.TrainTypes(words)
    l, r, n, c, vs = oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab)
    merges, _ = tt.merges_to_strs(l, r, n)
    return merges, vs


def test_fastbpe_pan_tadeusz_golden(pre_tokenize):
    lines = load_golden("pan_tadeusz.json.gz")
    gold = load_golden("pan_tadeusz.tokens.json.gz")
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    words = [pre_tokenize(l) for l in lines]
    arena, off = P.pack_words([w for ws in words for w in ws])
    ids, toff = oracle.bpe_encode(tab, arena, off)
    got = _split_by_offsets(tab.tokens_to_strs(ids), toff, [len(ws) for ws in words])
    assert got == gold["FastBPE"]
    assert len(ids) == 11117
    # NaiveBPE (in-order replay) on a prefix: it is O(merges) per word
    nb = sum(len(ws) for ws in words[:40])
    ids2, toff2 = oracle.bpe_encode(tab, *P.pack_words([w for ws in words for w in ws][:nb]), naive=True)
    assert tab.tokens_to_strs(ids2) == [t for l in gold["NaiveBPE"][:40] for t in l]


def test_fastwp_pan_tadeusz_golden(pre_tokenize):
    lines = load_golden("pan_tadeusz.json.gz")
    gold = load_golden("pan_tadeusz.tokens.json.gz")
    tab = P.WpTables(load_golden("pretrained_wp_vocab.json.gz"))
    alnum, space = P.unicode_class_bitmaps()
    trie = oracle.WpTrie(tab, alnum)
    st = trie.stats()
    # 50,172 nodes (+ root_p here), 50,171 edges, 50,277 pops, 122 links to root_p
    assert (st["nodes"], st["edges"], st["pops"], st["root_p_links"]) == (50173, 50171, 50277, 122)
    # whole lines (the reference's own input form) ...
    ids, toff, h6 = trie.encode(*P.pack_words([l.lower() for l in lines]), space)
    assert h6 == 0 and len(ids) == 29164
    strs = tab.tokens_to_strs(ids)
    assert [strs[int(toff[i]):int(toff[i + 1])] for i in range(len(lines))] == gold["FastWordPiece"]
    # ... and whitespace-free chunks (the product's input form) give the same stream
    ids_c, _, h6c = trie.encode(*P.pack_words([c for l in lines for c in l.lower().split()]), space)
    assert h6c == 0 and np.array_equal(ids, ids_c)
    # NaiveWP.encode_word over BERT-pre-tokenized words
    idsn, _ = trie.naive_encode(*P.pack_words([w for l in lines for w in pre_tokenize(l)]))
    assert tab.tokens_to_strs(idsn) == [t for l in gold["NaiveWordPiece"] for t in l]


def test_bpe_train_readme_kat(pre_tokenize):
    kat = load_golden("kat_tests_resources.json")
    words = [w for s in kat["corpus"] for w in pre_tokenize(s)]
    merges, vs = _bpe_train_oracle(words, kat["max_vocab"])
    assert merges == [tuple(p) for p in kat["FastBPE"]] == [tuple(p) for p in kat["NaiveBPE"]]
    assert vs == 25
    # overlapping-pair counting + greedy replacement
    merges, vs = _bpe_train_oracle(pre_tokenize("aaa aaa b"), 10)
    assert merges == [("a", "a"), ("aa", "a")]


def test_readme_sentence(pre_tokenize):
    kat = load_golden("kat_tests_resources.json")
    sent = kat["readme_sentence"]
    tab = P.BpeTables([tuple(p) for p in kat["FastBPE"]])
    ids, _ = oracle.bpe_encode(tab, *P.pack_words(pre_tokenize(sent)))
    assert tab.tokens_to_strs(ids) == kat["readme_tokens"]["FastBPE"]
    wt = P.WpTables(kat["FastWordPiece"])
    alnum, space = P.unicode_class_bitmaps()
    trie = oracle.WpTrie(wt, alnum)
    ids, _, h6 = trie.encode(*P.pack_words([sent.lower()]), space)
    assert h6 == 0 and wt.tokens_to_strs(ids) == kat["readme_tokens"]["FastWordPiece"]
    assert "['UNK']" in kat["readme_tokens"]["FastWordPiece"]
    ids, _ = trie.naive_encode(*P.pack_words(pre_tokenize(sent)))
    assert wt.tokens_to_strs(ids) == kat["readme_tokens"]["NaiveWordPiece"]


def test_bpe_train_5k_survey_hash(pre_tokenize):
    corpus = load_golden("train-5K.json.gz")
    words = [w for s in corpus for w in pre_tokenize(s)]
    merges, vs = _bpe_train_oracle(words, 1000)
    assert len(merges) == 922 and vs == 1000
    h = hashlib.sha256(json.dumps(merges, ensure_ascii=False).encode()).hexdigest()
    assert h == "f5f4451432124d34d4b1803a7482deebb3ac78d88e891cd6e8f10f5ad8967a44"
    assert merges == [tuple(p) for p in load_golden("ref_bpe_train5k_v1000_merges.json.gz")]
    lines = load_golden("pan_tadeusz.json.gz")
    tab = P.BpeTables(merges)
    pw = [pre_tokenize(l) for l in lines]
    ids, toff = oracle.bpe_encode(tab, *P.pack_words([w for ws in pw for w in ws]))
    out = _split_by_offsets(tab.tokens_to_strs(ids), toff, [len(ws) for ws in pw])
    assert len(ids) == 16469
    assert hashlib.sha256(json.dumps(out, ensure_ascii=False).encode()).hexdigest() == \
        "59c4833e7556ced4d1786716d3dc8fa31ac45a4c8d116fa4e159c4ca6a33ae58"


def test_random_bpe_train(random_cases, pre_tokenize):
    for case in random_cases["bpe_train"]:
        words = [w for s in case["corpus"] for w in pre_tokenize(s)]
        merges, vs = _bpe_train_oracle(words, case["max_vocab"])
        assert merges == [tuple(p) for p in case["merges"]], case["corpus"]
        assert vs == case["vocab_size"]


def test_random_bpe_encode(random_cases):
    for case in random_cases["bpe_encode"]:
        tab = P.BpeTables([tuple(p) for p in case["merges"]])
        arena, off = P.pack_words(case["words"])
        for naive, key in ((False, "fast"), (True, "naive")):
            ids, toff = oracle.bpe_encode(tab, arena, off, naive=naive)
            strs = tab.tokens_to_strs(ids)
            got = [strs[int(toff[i]):int(toff[i + 1])] for i in range(len(case["words"]))]
            assert got == case[key]


def test_random_wp_encode(random_cases, pre_tokenize):
    alnum, space = P.unicode_class_bitmaps()
    n_checked = n_hang = 0
    for case in random_cases["wp_encode"]:
        tab = P.WpTables(case["vocab"])
        trie = oracle.WpTrie(tab, alnum)
        for text, fast, naive in zip(case["texts"], case["fast"], case["naive"]):
            ids, _, h6 = trie.encode(*P.pack_words([text.lower()]), space)
            if fast is None:            # the reference does not terminate here (H6)
                assert h6 > 0
                n_hang += 1
            else:
                assert h6 == 0
                assert tab.tokens_to_strs(ids) == fast, (case["vocab"], text)
                ids_c, _, _ = trie.encode(*P.pack_words(text.lower().split()), space)
                assert np.array_equal(ids, ids_c)
                n_checked += 1
            if naive is not None:
                ids, _ = trie.naive_encode(*P.pack_words(pre_tokenize(text)))
                assert tab.tokens_to_strs(ids) == naive
    assert n_checked > 800


def _wp_train_oracle(words, max_vocab):
    from subword_tokenizers_b200.packing import WpTrainTypes
    tt = WpTrainTypes(words)
    l, r, n, vs = oracle.wp_train(tt.syms, tt.off, tt.freq, tt.init_cps, tt.init_off, max_vocab)
    return tt.vocab_from_merges(l, r, n), vs


def test_random_wp_train(random_cases, pre_tokenize):
    for case in random_cases["wp_train"]:
        words = [w for s in case["corpus"] for w in pre_tokenize(s)]
        vocab, vs = _wp_train_oracle(words, case["max_vocab"])
        assert sorted(vocab) == case["vocab"], case["corpus"]
        assert vs == len(case["vocab"])


def test_wp_train_readme_kat_and_5k(pre_tokenize):
    kat = load_golden("kat_tests_resources.json")
    words = [w for s in kat["corpus"] for w in pre_tokenize(s)]
    vocab, _ = _wp_train_oracle(words, 25)
    assert set(vocab) == set(kat["FastWordPiece"]) == set(kat["NaiveWordPiece"])
    corpus = load_golden("train-5K.json.gz")
    words = [w for s in corpus for w in pre_tokenize(s)]
    vocab, vs = _wp_train_oracle(words, 1000)
    assert sorted(vocab) == load_golden("ref_wp_train5k_v1000_vocab.json.gz")
