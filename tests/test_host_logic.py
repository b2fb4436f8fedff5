"""Host-side logic that needs no GPU: packing, the Naive* host classes against the reference's golden
outputs, error behaviour and the on-disk layout."""
import json

import numpy as np
import pytest

from conftest import load_golden
from subword_tokenizers_b200 import FastBPE, FastWP, NaiveBPE, NaiveWP, packing as P
from subword_tokenizers_b200.device import shard_types


def test_pack_words_roundtrip():
    words = ["", "zażółć", "a", "\U0001F600x", "ab" * 40]
    arena, off = P.pack_words(words)
    assert off[0] == 0 and off[-1] == len(arena)
    back = [P.decode_utf8(arena[int(off[i]):int(off[i + 1])].tobytes()) for i in range(len(words))]
    assert back == words


def test_bpe_tables_canonical_ids_and_last_rank():
    tab = P.BpeTables([("a", "b"), ("ab", "c"), ("a", "bc"), ("a", "b")])
    assert tab.id_to_str[tab.new[1]] == "abc" and tab.new[1] == tab.new[2]        # one id per distinct string (H3)
    assert tab.new[0] == tab.new[3]
    assert tab.token_to_str(P.BPE_EMPTY_TOKEN) == ""
    assert tab.token_to_str(((P.BPE_UNKNOWN_CP | ord("ż")) << 1) | 1) == "##ż"


def test_unicode_class_bitmaps_match_python():
    alnum, space = P.unicode_class_bitmaps()
    for cp in list(range(0, 0x300)) + [0x1C, 0x1F, 0x85, 0xA0, 0x2028, 0x3000, 0x20AC, 0x1F600, 0x10FFFF]:
        ch = chr(cp)
        assert bool((alnum[cp >> 3] >> (cp & 7)) & 1) == ch.isalnum()
        assert bool((space[cp >> 3] >> (cp & 7)) & 1) == ch.isspace()


def test_train_types_first_occurrence_order():
    tt = P.TrainTypes(["b", "a", "b", "ca", "a", "b"])
    assert tt.types == ["b", "a", "ca"] and tt.freq.tolist() == [3, 2, 1]
    assert tt.alphabet == ["a", "b", "c"] and tt.syms.tolist() == [1, 0, 2, 0] and tt.off.tolist() == [0, 1, 2, 4]
    merges, strs = tt.merges_to_strs(np.array([2]), np.array([0]), np.array([3]))
    assert merges == [("c", "a")] and strs[3] == "ca"


def test_shard_types_contiguous_and_complete():
    off = np.array([0, 5, 6, 20, 21, 22, 40], dtype=np.uint64)
    for world in (1, 2, 3, 4, 8):
        shards = shard_types(off, world)
        assert shards[0][0] == 0 and shards[-1][1] == 6
        assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))


def test_type_errors_match_reference_messages(hf_tokenizer):
    for cls in (NaiveBPE, FastBPE):
        t = cls(hf_tokenizer)
        with pytest.raises(TypeError, match="Corpus must be a list of strings."):
            t.train("not a list", 10)
        with pytest.raises(TypeError, match="Maximum vocabulary size must be an integer."):
            t.train(["a"], 10.5)
        with pytest.raises(TypeError):
            t.tokenize(3)
    for cls in (NaiveWP, FastWP):
        t = cls(hf_tokenizer)
        with pytest.raises(TypeError, match="corpus must be a list of strings."):
            t.train(["a", 3], 10)
        with pytest.raises(TypeError, match="max_vocab must be an int."):
            t.train(["a"], "10")
        with pytest.raises(TypeError):
            t.tokenize(None)
    with pytest.raises(AttributeError):
        FastWP(hf_tokenizer).tokenize("trie not built yet")


def test_resources_layout_and_missing_file(hf_tokenizer, tmp_path):
    nb = NaiveBPE(hf_tokenizer)
    nb.merges_list = [("a", "ł"), ("ał", "b")]
    nb.save_resources(str(tmp_path / "NaiveBPE"))
    raw = open(tmp_path / "NaiveBPE" / "merges.json", encoding="utf-8").read()
    assert raw == '[["a", "ł"], ["ał", "b"]]'                       # json.dump(..., ensure_ascii=False), no indent
    nb2 = NaiveBPE(hf_tokenizer)
    nb2.load_resources(str(tmp_path / "NaiveBPE"))
    assert nb2.merges_list == [("a", "ł"), ("ał", "b")]
    nb2.load_resources(str(tmp_path / "does-not-exist"))             # silently ignored
    assert nb2.merges_list == [("a", "ł"), ("ał", "b")]
    nw = NaiveWP(hf_tokenizer)
    nw.vocab = {"a", "##b", "ł"}
    nw.save_resources(str(tmp_path / "NaiveWordPiece"))
    assert sorted(json.load(open(tmp_path / "NaiveWordPiece" / "vocab.json", encoding="utf-8"))) == ["##b", "a", "ł"]
    nw2 = NaiveWP(hf_tokenizer)
    nw2.load_resources(str(tmp_path / "NaiveWordPiece"))
    assert nw2.vocab == {"a", "##b", "ł"}


def test_naive_bpe_host_encoder_matches_reference(hf_tokenizer, random_cases):
    nb = NaiveBPE(hf_tokenizer)
    for case in random_cases["bpe_encode"]:
        nb.merges_list = [tuple(p) for p in case["merges"]]
        assert [nb.encode_word(w) for w in case["words"]] == case["naive"]
    gold = load_golden("pan_tadeusz.tokens.json.gz")["NaiveBPE"]
    nb.merges_list = [tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")]
    lines = load_golden("pan_tadeusz.json.gz")
    # encode_word is the host replay (bpe.py:114-132); tokenize() runs on the GPU (tests/test_gpu_parity.py)
    assert [[t for w in nb._pre_tokenized_words([l]) for t in nb.encode_word(w)] for l in lines[:3]] == gold[:3]
    assert nb.encode_word("") == []


def test_naive_wp_host_encoder_matches_reference(hf_tokenizer, random_cases):
    """NaiveWP.encode_word is host code (wordpiece.py:131-158); NaiveWP.tokenize / train run on the GPU
    (tests/test_gpu_parity.py)."""
    nw = NaiveWP(hf_tokenizer)

    def host_tokenize(text):
        return [t for w in nw._pre_tokenized_words([text]) for t in nw.encode_word(w)]
    for case in random_cases["wp_encode"]:
        nw.vocab = set(case["vocab"])
        for text, naive in zip(case["texts"], case["naive"]):
            if naive is not None:
                assert host_tokenize(text) == naive
    nw.vocab = set(load_golden("pretrained_wp_vocab.json.gz"))
    lines = load_golden("pan_tadeusz.json.gz")
    assert [host_tokenize(l) for l in lines] == load_golden("pan_tadeusz.tokens.json.gz")["NaiveWordPiece"]


def test_pretok_tables_come_from_the_running_interpreter():
    """Tables of the device pre-tokenizer (packing.PretokTables): generated from CPython's own str.lower / str.isspace."""
    from subword_tokenizers_b200 import packing as P
    t = P.PretokTables.get()
    assert {c for c in range(0x110000) if chr(c).isspace()} == set(P.PY_SPACE_CPS)
    for ch in "ŁÉΩЖ\U00010400ẞ\u212a\u2126":                      # (ASCII is folded by SWAR arithmetic in the kernel, not by the table)
        assert int(t.lower_map[ord(ch)]) == ord(ch.lower())
    e = int(t.lower_map[0x130])
    assert e & P.LOWER_MULTI and [int(x) for x in t.multi[(e & 0xFFFFF):(e & 0xFFFFF) + 3]] == [2, 0x69, 0x307]
    assert int(t.lower_map[0x3A3]) == P.LOWER_SIGMA
    assert all(int(t.lower_map[c]) == c for c in (0x142, 0x4E2D, 0x20AC))
    cased, ign = (np.unpackbits(b, bitorder="little") for b in t.sigma_bitmaps())
    assert cased[ord("a")] and cased[ord("Σ")] and not cased[ord("1")] and not cased[ord(" ")]
    assert ign[ord("'")] and ign[0x301] and ign[ord(".")] and not ign[ord("a")] and not ign[ord(" ")] and not ign[ord("-")]
    assert cased[0x2B0] and ign[0x2B0]                                   # MODIFIER LETTER SMALL H: cased and case-ignorable


def test_bert_class_table_matches_the_tokenizers_library(hf_tokenizer):
    """data/bert_pretok_classes.json was probed from the Rust BertPreTokenizer; it must agree with the library installed here."""
    from subword_tokenizers_b200 import packing as P
    pre = hf_tokenizer.backend_tokenizer.pre_tokenizer
    assert P.bert_pretokenizer_matches(pre)
    cls = P.load_bert_classes()
    assert sum(b - a + 1 for a, b in cls["punct"]) == 726 and sum(b - a + 1 for a, b in cls["space"]) == 25
    m = P.PretokTables.get().bert_lower_map()
    assert m[ord("«")] & P.LOWER_PUNCT and m[0x3001] & P.LOWER_PUNCT and not (m[ord("é")] & P.LOWER_PUNCT)

    class Other:                                                           # any other pre-tokenizer keeps the host path
        def pre_tokenize_str(self, s):
            return [(w, (0, 0)) for w in s.split()]
    assert not P.bert_pretokenizer_matches(Other())


def test_encode_texts_splits_large_batches_on_the_host():
    """_Encoder.encode_texts cuts a batch that exceeds BATCH_TEXT_BYTES into several device passes and stitches ids / per-text
    cuts back together (host logic; the device pass is replaced by a stand-in that emits one id per whitespace word)."""
    from subword_tokenizers_b200.device import _Encoder

    class Stub(_Encoder):
        calls = 0

        def __init__(self):                      # no device handle
            pass

        def _encode_texts_once(self, enc):
            Stub.calls += 1
            counts = [len(b.split()) for b in enc]
            ids = np.array([len(w) for b in enc for w in b.split()], dtype=np.uint32)
            return ids, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)

        def __del__(self):
            pass

    texts = ["a bb ccc", "", "dddd e", "ff", "   ", "g hh iii jjjj"] * 7
    s = Stub()
    ids_1, cut_1 = s.encode_texts(texts)
    assert Stub.calls == 1 and cut_1[-1] == len(ids_1) and len(cut_1) == len(texts) + 1
    s.BATCH_TEXT_BYTES = 20
    Stub.calls = 0
    ids_n, cut_n = s.encode_texts(texts)
    assert Stub.calls > 5
    assert np.array_equal(ids_1, ids_n) and np.array_equal(cut_1, cut_n)
    for k, t in enumerate(texts):
        assert ids_n[cut_n[k]:cut_n[k + 1]].tolist() == [len(w) for w in t.split()]


def test_token_id_to_string_tables():
    """Vectorised id -> string materialisation (packing.*Tables.tokens_to_strs) against the scalar definitions of include/swt.h."""
    bt = P.BpeTables([("a", "b"), ("ab", "c"), ("x", "y")])
    sym = {s: i for i, s in enumerate(bt.id_to_str)}
    toks = np.array([sym["abc"] << 1, (sym["xy"] << 1) | 1, ((P.BPE_UNKNOWN_CP | ord("ż")) << 1) & 0xFFFFFFFF | 1,
                     (P.BPE_UNKNOWN_CP | 0x1F600) << 1 & 0xFFFFFFFF, P.BPE_EMPTY_TOKEN, (sym["a"] << 1) | 1], dtype=np.uint32)
    want = [bt.token_to_str(int(t)) for t in toks]
    assert want[:2] == ["abc", "##xy"] and want[4] == "" and want[5] == "##a"
    assert bt.tokens_to_strs(toks) == want
    assert bt.tokens_to_strs(toks[:2]) == want[:2] and bt.tokens_to_strs([]) == []
    wt = P.WpTables(["b", "##a", "a"])
    assert wt.tokens_to_strs(np.array([0, 1, 2, 3, 4], dtype=np.uint32)) == ["##a", "a", "b", "['UNK']", "[UNK]"]
    assert wt.tokens_to_strs([]) == []


def test_config5_fixture_and_inputs_are_consistent():
    """The config-5 fixture (tests/golden/make_config5_fixture.py, unmodified reference) belongs to the committed 50 K models and to
    the deterministic adversarial inputs of tests/golden/config5_inputs.py."""
    import os
    import sys
    import bench_data as BD
    from conftest import ROOT, load_golden
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from config5_inputs import config5_sentences
    fixture = load_golden("config5_fixture.json.gz")
    merges, vocab = load_golden("config5_bpe_merges.json.gz"), load_golden("config5_wp_vocab.json.gz")
    assert len(vocab) == 50_000 and len(set(vocab)) == 50_000 and 49_000 < len(merges) <= 50_000
    mat, lens = BD.synth_type_table(300_000, 9)
    arena, off = BD.table_to_utf8(mat, lens)
    types = [arena[int(off[k]):int(off[k + 1])].tobytes().decode() for k in range(2000)]
    s1, s2 = config5_sentences(types, set(vocab)), config5_sentences(types, set(vocab))
    assert s1 == s2 and len(s1) == fixture["n_sentences"]
    assert set(fixture["equivalence"]) == {"NaiveBPE/FastBPE", "NaiveWP/FastWP", "FastBPE/FastWP"}
    assert fixture["equivalence"]["NaiveBPE/FastBPE"][2] == 100.0          # the trained merge list is in training order: Naive == Fast
    assert fixture["n_tokens"]["NaiveBPE"] == fixture["n_tokens"]["FastBPE"] and fixture["n_tokens"]["NaiveWP"] == fixture["n_tokens"]["FastWP"]
