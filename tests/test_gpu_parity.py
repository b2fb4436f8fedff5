"""GPU parity tests proper: every call goes through the C ABI (libswt.so) on cuda:0 and is compared with
the CPU oracle on the same inputs and with the committed golden vectors produced by the reference."""
import hashlib
import json

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from subword_tokenizers_b200 import packing
    return packing


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    from subword_tokenizers_b200 import device
    return device


def _split(strs, tok_off, counts):
    out, wi = [], 0
    for n in counts:
        out.append(strs[int(tok_off[wi]):int(tok_off[wi + n])])
        wi += n
    return out


# ------------------------------------------------------------------------------------------------ HP-1
def test_fastbpe_pan_tadeusz_golden(P, dev, pre_tokenize):
    import oracle
    lines = load_golden("pan_tadeusz.json.gz")
    gold = load_golden("pan_tadeusz.tokens.json.gz")["FastBPE"]
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    enc = dev.BpeEncoder(tab)
    words = [pre_tokenize(l) for l in lines]
    flat = [w for ws in words for w in ws]
    ids, tok_off, _ = enc.encode_words(flat)
    assert _split(tab.tokens_to_strs(ids), tok_off, [len(ws) for ws in words]) == gold
    o_ids, o_off = oracle.bpe_encode(tab, *P.pack_words(flat))
    assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off)


def test_fastbpe_random_cases(P, dev, random_cases):
    import oracle
    for case in random_cases["bpe_encode"]:
        tab = P.BpeTables([tuple(p) for p in case["merges"]])
        enc = dev.BpeEncoder(tab)
        ids, tok_off, _ = enc.encode_words(case["words"])
        strs = tab.tokens_to_strs(ids)
        assert [strs[int(tok_off[i]):int(tok_off[i + 1])] for i in range(len(case["words"]))] == case["fast"]
        o_ids, o_off = oracle.bpe_encode(tab, *P.pack_words(case["words"]))
        assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off)
        enc.close()


def _adversarial_words(rng, alphabet):
    words = ["", "a", "ab"]
    for n in (31, 32, 33, 255, 256, 257, 4096, 65536):
        words.append("a" * n)                                                   # runs: overlap semantics (H2)
        words.append("".join(rng.choice(list(alphabet), size=n)))               # long random words
    words += ["".join(rng.choice(list(alphabet), size=int(rng.integers(1, 40)))) for _ in range(3000)]
    words += ["ż" * 17, "€uro", "\U0001F600ab", "ab\U0001F600", "q" * 40]          # non-ASCII, unknown chars
    return words


def test_fastbpe_adversarial_vs_oracle(P, dev):
    import oracle
    rng = np.random.default_rng(5)
    merges = [tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")]
    merges += [("a", "a"), ("aa", "aa"), ("aaaa", "aaaa"), ("aaaaaaaa", "aaaaaaaa"), ("b", "a"), ("a", "b")]
    tab = P.BpeTables(merges)
    enc = dev.BpeEncoder(tab)
    words = _adversarial_words(rng, "abcdeiknorstwyzł")
    ids, tok_off, _ = enc.encode_words(words)
    o_ids, o_off = oracle.bpe_encode(tab, *P.pack_words(words))
    assert np.array_equal(tok_off.astype(np.uint64), o_off)
    assert np.array_equal(ids, o_ids)
    # untrainable-order list: FastBPE semantics, not NaiveBPE's
    tab2 = P.BpeTables([("ab", "c"), ("a", "b")])
    enc2 = dev.BpeEncoder(tab2)
    ids2, _, _ = enc2.encode_words(["abc"])
    assert tab2.tokens_to_strs(ids2) == ["abc"]


def test_bpe_empty_inputs(P, dev):
    tab = P.BpeTables([("a", "b")])
    enc = dev.BpeEncoder(tab)
    ids, tok_off, _ = enc.encode_words([])
    assert len(ids) == 0 and list(tok_off) == [0]
    ids, tok_off, _ = enc.encode_words(["", "ab", ""])
    assert tab.tokens_to_strs(ids) == ["", "ab", ""] and list(tok_off) == [0, 1, 2, 3]


# ------------------------------------------------------------------------------------------------ HP-2
def _wp_encoder(P, dev, vocab):
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    tab = P.WpTables(vocab)
    return tab, dev.WpEncoder(tab, naive_wp_encode_ids("##", tab))


def test_fastwp_pan_tadeusz_golden(P, dev):
    import oracle
    lines = load_golden("pan_tadeusz.json.gz")
    gold = load_golden("pan_tadeusz.tokens.json.gz")["FastWordPiece"]
    tab, enc = _wp_encoder(P, dev, load_golden("pretrained_wp_vocab.json.gz"))
    st = enc.stats()
    assert (st["nodes"], st["edges"], st["pops"], st["root_p_links"]) == (50173, 50171, 50277, 122)
    chunks = [l.lower().split() for l in lines]
    flat = [c for cs in chunks for c in cs]
    ids, tok_off, h6 = enc.encode_words(flat)
    assert h6 == 0
    assert _split(tab.tokens_to_strs(ids), tok_off, [len(cs) for cs in chunks]) == gold
    alnum, space = P.unicode_class_bitmaps()
    o_ids, o_off, _ = oracle.WpTrie(tab, alnum).encode(*P.pack_words(flat), space)
    assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off)


def test_fastwp_random_cases(P, dev, random_cases):
    import oracle
    alnum, space = P.unicode_class_bitmaps()
    n_ref = 0
    for case in random_cases["wp_encode"]:
        try:
            tab, enc = _wp_encoder(P, dev, case["vocab"])
        except NotImplementedError:
            continue
        trie = oracle.WpTrie(tab, alnum)
        for text, fast in zip(case["texts"], case["fast"]):
            chunks = text.lower().split()
            ids, tok_off, h6 = enc.encode_words(chunks)
            o_ids, o_off, o_h6 = trie.encode(*P.pack_words(chunks), space)
            assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off) and h6 == o_h6
            if fast is not None:                      # the reference terminated on this text
                assert h6 == 0 and tab.tokens_to_strs(ids) == fast
                n_ref += 1
        enc.close()
    assert n_ref > 500


def test_fastwp_adversarial_vs_oracle(P, dev):
    import oracle
    rng = np.random.default_rng(7)
    vocab = load_golden("pretrained_wp_vocab.json.gz")
    tab, enc = _wp_encoder(P, dev, vocab)
    alnum, space = P.unicode_class_bitmaps()
    trie = oracle.WpTrie(tab, alnum)
    alphabet = list("abcdeiknorstwyzł.,-!")
    chunks = ["", "a", "##", "##a", "a##", ".", "...", "a.b", "€", "a€b", "\U0001F600", "x" * 33]
    for n in (31, 32, 33, 255, 256, 257, 4096, 65536):
        chunks.append("a" * n)
        chunks.append("".join(rng.choice(alphabet, size=n)))
    chunks += ["".join(rng.choice(alphabet, size=int(rng.integers(1, 40)))) for _ in range(5000)]
    ids, tok_off, h6 = enc.encode_words(chunks)
    o_ids, o_off, o_h6 = trie.encode(*P.pack_words(chunks), space)
    assert np.array_equal(tok_off.astype(np.uint64), o_off)
    assert np.array_equal(ids, o_ids) and h6 == o_h6 and h6 > 0


# ------------------------------------------------------------------------------------------------ HP-3
def _train_gpu(dev, P, words, max_vocab, **kw):
    tt = P.TrainTypes(words)
    max_len = int(np.diff(tt.off.astype(np.int64)).max()) if tt.n_types else 1
    eng = dev.CudaTrainEngine(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab, tt.n_alpha, max_len, 0, 0, 1, **kw)
    l, r, n, c, state = dev.run_training_loop(eng, 1, steps_per_sync=kw.get("record_cap", 64))
    return tt, l, r, n, c, state, eng


def _train_oracle(P, words, max_vocab):
    import oracle
    tt = P.TrainTypes(words)
    return oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab)


def test_bpe_train_kats(P, dev, pre_tokenize):
    kat = load_golden("kat_tests_resources.json")
    words = [w for s in kat["corpus"] for w in pre_tokenize(s)]
    tt, l, r, n, c, state, _ = _train_gpu(dev, P, words, kat["max_vocab"])
    merges, _ = tt.merges_to_strs(l, r, n)
    assert merges == [tuple(p) for p in kat["FastBPE"]]
    assert state["vocab_size"] == 25 and state["halt"] == 1
    tt, l, r, n, c, state, _ = _train_gpu(dev, P, pre_tokenize("aaa aaa b"), 10)
    assert tt.merges_to_strs(l, r, n)[0] == [("a", "a"), ("aa", "a")]
    assert state["halt"] == 2                                     # no pairs left (bpe.py:98-99)


def test_bpe_train_random_cases(P, dev, random_cases, pre_tokenize):
    for case in random_cases["bpe_train"]:
        words = [w for s in case["corpus"] for w in pre_tokenize(s)]
        tt, l, r, n, c, state, _ = _train_gpu(dev, P, words, case["max_vocab"], record_cap=7)
        merges, _ = tt.merges_to_strs(l, r, n)
        assert merges == [tuple(p) for p in case["merges"]], case["corpus"]
        assert state["vocab_size"] == case["vocab_size"]
        ol, orr, on, oc, ovs = _train_oracle(P, words, case["max_vocab"])
        assert np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on) and np.array_equal(c, oc)


def test_bpe_train_5k_config1(P, dev, pre_tokenize):
    corpus = load_golden("train-5K.json.gz")
    words = [w for s in corpus for w in pre_tokenize(s)]
    tt, l, r, n, c, state, eng = _train_gpu(dev, P, words, 1000, record_cap=256)
    merges, strs = tt.merges_to_strs(l, r, n)
    assert len(merges) == 922 and state["vocab_size"] == 1000
    assert merges == [tuple(p) for p in load_golden("ref_bpe_train5k_v1000_merges.json.gz")]
    h = hashlib.sha256(json.dumps(merges, ensure_ascii=False).encode()).hexdigest()
    assert h == "f5f4451432124d34d4b1803a7482deebb3ac78d88e891cd6e8f10f5ad8967a44"
    ol, orr, on, oc, _ = _train_oracle(P, words, 1000)
    assert np.array_equal(c, oc)                                  # chosen pair counts agree step by step
    # the device word table equals a replay of the merges (merge-apply parity)
    syms, lens = eng.read_corpus()
    total = 0
    for k in range(0, tt.n_types, 997):
        word = list(tt.types[k])
        for a, b in merges:
            out, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(word[i]); i += 1
            word = out
        s = int(tt.off[k])
        assert [strs[j] for j in syms[s:s + int(lens[k])]] == word
        total += 1
    assert total > 20


def test_bpe_train_table_growth_and_ties(P, dev):
    """Tiny pair table -> several rehashes; heavy ties (every type has freq 1)."""
    rng = np.random.default_rng(11)
    alphabet = list("abcdefghijklmnopqrstuvwxyz")
    words = ["".join(rng.choice(alphabet, size=int(rng.integers(2, 12)))) for _ in range(4000)]
    tt, l, r, n, c, state, _ = _train_gpu(dev, P, words, 600, record_cap=64, table_cap=1024)
    ol, orr, on, oc, ovs = _train_oracle(P, words, 600)
    assert np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on) and np.array_equal(c, oc)
    assert state["vocab_size"] == ovs


def test_bpe_train_long_word_runs(P, dev):
    words = ["a" * 5000, "ab" * 3000, "a" * 17, "b", "aab" * 1000] * 2 + ["ba" * 7]
    tt, l, r, n, c, state, _ = _train_gpu(dev, P, words, 40, record_cap=16)
    ol, orr, on, oc, ovs = _train_oracle(P, words, 40)
    assert np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on) and np.array_equal(c, oc)


# ------------------------------------------------------------------------------------------------ NaiveWP.train
def _wp_train_gpu(dev, P, words, max_vocab, **kw):
    from subword_tokenizers_b200 import _lib
    tt = P.WpTrainTypes(words)
    max_len = int(np.diff(tt.off.astype(np.int64)).max()) if len(tt.types) else 1
    eng = dev.CudaTrainEngine(tt.syms, tt.off, tt.freq, len(tt.init_syms), max_vocab, len(tt.init_syms), max_len + 2, 0, 0, 1,
                              mode=_lib.TRAIN_WP, init_cps=tt.init_cps, init_off=tt.init_off, **kw)
    l, r, n, c, state = dev.run_training_loop(eng, 1, steps_per_sync=kw.get("record_cap", 64))
    return tt, l, r, n, state


def test_wp_train_random_cases_and_kat(P, dev, random_cases, pre_tokenize):
    import oracle
    for case in random_cases["wp_train"]:
        words = [w for s in case["corpus"] for w in pre_tokenize(s)]
        tt, l, r, n, state = _wp_train_gpu(dev, P, words, case["max_vocab"], record_cap=5)
        assert sorted(tt.vocab_from_merges(l, r, n)) == case["vocab"], case["corpus"]
        assert state["vocab_size"] == len(case["vocab"])
        ol, orr, on, ovs = oracle.wp_train(tt.syms, tt.off, tt.freq, tt.init_cps, tt.init_off, case["max_vocab"])
        assert np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on)
    kat = load_golden("kat_tests_resources.json")
    words = [w for s in kat["corpus"] for w in pre_tokenize(s)]
    tt, l, r, n, state = _wp_train_gpu(dev, P, words, kat["max_vocab"])
    assert set(tt.vocab_from_merges(l, r, n)) == set(kat["NaiveWordPiece"])


def test_wp_train_5k_matches_reference(P, dev, pre_tokenize):
    import oracle
    words = [w for s in load_golden("train-5K.json.gz") for w in pre_tokenize(s)]
    tt, l, r, n, state = _wp_train_gpu(dev, P, words, 8000, record_cap=512)
    vocab = tt.vocab_from_merges(l, r, n)
    assert sorted(vocab) == load_golden("ref_wp_train5k_v8000_vocab.json.gz")          # the reference's own 40-minute run
    ol, orr, on, ovs = oracle.wp_train(tt.syms, tt.off, tt.freq, tt.init_cps, tt.init_off, 8000)
    assert np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on) and state["vocab_size"] == ovs
    # a prefix of the same run is the 1000-entry vocabulary
    tt2, l2, r2, n2, _ = _wp_train_gpu(dev, P, words, 1000, record_cap=512)
    assert sorted(tt2.vocab_from_merges(l2, r2, n2)) == load_golden("ref_wp_train5k_v1000_vocab.json.gz")


# ------------------------------------------------------------------------------------------------ classes
def test_classes_end_to_end(hf_tokenizer, tmp_path):
    from subword_tokenizers_b200 import FastBPE, FastWP, NaiveBPE
    kat = load_golden("kat_tests_resources.json")
    fb = FastBPE(hf_tokenizer)
    fb.train(kat["corpus"], kat["max_vocab"])
    assert fb.merges_list == [tuple(p) for p in kat["FastBPE"]] and len(fb.vocab) == 25
    assert fb.tokenize(kat["readme_sentence"]) == kat["readme_tokens"]["FastBPE"]
    assert fb.corpus_as_symbols[0] == (["this"], 1)
    fb.save_resources(str(tmp_path / "FastBPE"))
    assert json.load(open(tmp_path / "FastBPE" / "merges.json")) == kat["FastBPE"]
    fb2 = FastBPE(hf_tokenizer)
    fb2.load_resources(str(tmp_path / "FastBPE"))
    assert fb2.tokenize(kat["readme_sentence"]) == kat["readme_tokens"]["FastBPE"]
    assert fb2.encode_word("") == [""] and fb2.encode_word("x") == ["x"]
    nb = NaiveBPE(hf_tokenizer)
    nb.train(kat["corpus"], kat["max_vocab"])
    assert nb.merges_list == fb.merges_list
    assert nb.tokenize(kat["readme_sentence"]) == kat["readme_tokens"]["NaiveBPE"]
    fw = FastWP(hf_tokenizer)
    with pytest.raises(AttributeError):
        fw.tokenize("no trie yet")
    fw.train(kat["corpus"], kat["max_vocab"])
    assert fw.vocab == set(kat["FastWordPiece"])
    assert fw.corpus_as_symbols[0][1] == 1 and "".join(s.replace("##", "") for s in fw.corpus_as_symbols[0][0]) == "this"
    assert fw.tokenize(kat["readme_sentence"]) == kat["readme_tokens"]["FastWordPiece"]
    assert fw.tokenize_batch([kat["readme_sentence"], ""]) == [kat["readme_tokens"]["FastWordPiece"], []]


# ------------------------------------------------------------------------------------------------ host-buffer ABI
def test_encode_host_pipeline_matches_device_call(P, dev):
    import torch
    rng = np.random.default_rng(3)
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    wtab, wenc = _wp_encoder(P, dev, load_golden("pretrained_wp_vocab.json.gz"))
    benc = dev.BpeEncoder(tab)
    types = load_golden("pan_tadeusz.json.gz")
    words = [w for l in types for w in l.lower().split()]
    words = [words[i] for i in rng.integers(0, len(words), size=200_000)] + ["x" * 5000]
    arena, off = P.pack_words(words)
    off32 = off.astype(np.uint32)
    for enc in (benc, wenc):
        ids, tok_off, h6 = enc.encode_packed(arena, off32)
        h_arena = torch.from_numpy(arena).pin_memory()
        h_off = torch.from_numpy(off32.view(np.int32)).pin_memory()
        h_ids = torch.empty(len(arena) + len(words) + 16, dtype=torch.int32).pin_memory()
        h_tok = torch.empty(len(words) + 1, dtype=torch.int32).pin_memory()
        nt, h6b = enc.encode_host(h_arena, h_off, h_ids, h_tok, batch_bytes=1 << 17)    # many small batches
        assert nt == len(ids) and h6b == h6
        assert np.array_equal(h_ids.numpy()[:nt].view(np.uint32), ids)
        assert np.array_equal(h_tok.numpy().view(np.uint32), tok_off)


# ------------------------------------------------------------------------------------------------ scale properties
def test_large_stream_properties(P, dev):
    """At sizes the oracle cannot replay in seconds: size-independent properties.
    (1) the token stream of a word stream drawn from a type list equals the gather of the per-type
    encodings (each word is encoded independently); (2) token offsets are sorted and end at n_tokens;
    (3) decoding the ids reproduces the input bytes for BPE (tokens concatenate to the word)."""
    rng = np.random.default_rng(1)
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    wtab, wenc = _wp_encoder(P, dev, load_golden("pretrained_wp_vocab.json.gz"))
    benc = dev.BpeEncoder(tab)
    types = sorted({w for l in load_golden("train-5K.json.gz")[:2000] for w in l.lower().split()})
    n = 3_000_000
    zipf = 1.0 / np.arange(1, len(types) + 1)
    draw = rng.choice(len(types), size=n, p=zipf / zipf.sum())
    t_arena, t_off = P.pack_words(types)
    lens = np.diff(t_off.astype(np.int64))
    w_len = lens[draw]
    off = np.zeros(n + 1, dtype=np.int64); np.cumsum(w_len, out=off[1:])
    idx = np.repeat(t_off[:-1].astype(np.int64)[draw] - off[:-1], w_len) + np.arange(off[-1])
    arena = t_arena[idx]
    for enc in (benc, wenc):
        t_ids, t_tok_off, _ = enc.encode_packed(t_arena, t_off.astype(np.uint32))
        ids, tok_off, _ = enc.encode_packed(arena, off.astype(np.uint32))
        tl = np.diff(t_tok_off.astype(np.int64))
        assert np.array_equal(np.diff(tok_off.astype(np.int64)), tl[draw])
        assert int(tok_off[-1]) == len(ids)
        gidx = np.repeat(t_tok_off[:-1].astype(np.int64)[draw] - tok_off[:-1].astype(np.int64), tl[draw]) + np.arange(len(ids))
        assert np.array_equal(ids, t_ids[gidx])


# ------------------------------------------------------------------------------------------------ memo / capacity edge cases
def test_all_distinct_words_overflow_the_memo(P, dev):
    """No repetition at all: the memo (n_words/4 slots) fills up, later words take the direct / recompute path."""
    import oracle
    rng = np.random.default_rng(21)
    alphabet = list("abcdeiknorstwyzłąę.,")
    words = list({"".join(rng.choice(alphabet, size=int(rng.integers(1, 30)))) for _ in range(120_000)})
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    wtab, wenc = _wp_encoder(P, dev, load_golden("pretrained_wp_vocab.json.gz"))
    benc = dev.BpeEncoder(tab)
    arena, off = P.pack_words(words)
    alnum, space = P.unicode_class_bitmaps()
    ids, tok_off, _ = benc.encode_packed(arena, off.astype(np.uint32))
    o_ids, o_off = oracle.bpe_encode(tab, arena, off)
    assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off)
    ids, tok_off, h6 = wenc.encode_packed(arena, off.astype(np.uint32))
    o_ids, o_off, o_h6 = oracle.WpTrie(wtab, alnum).encode(arena, off, space)
    assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off) and h6 == o_h6


def test_words_sharing_a_15_byte_prefix(P, dev):
    """The 128-bit memo key is the first 15 bytes + length; longer words with equal prefix and length must not alias."""
    import oracle
    base = "abcdefghijklmno"                                    # 15 bytes
    words = [base + s for s in ("p", "q", "pq", "qp", "pqrstuvwxyzabcdef"[:17], "pqrstuvwxyzabcdeg"[:17])] * 50
    words += [base, base + "p"] * 50
    tab = P.BpeTables([("a", "b"), ("ab", "c"), ("o", "p"), ("o", "q"), ("p", "q")])
    benc = dev.BpeEncoder(tab)
    arena, off = P.pack_words(words)
    ids, tok_off, _ = benc.encode_packed(arena, off.astype(np.uint32))
    o_ids, o_off = oracle.bpe_encode(tab, arena, off)
    assert np.array_equal(ids, o_ids) and np.array_equal(tok_off.astype(np.uint64), o_off)


def test_output_capacity_is_checked(P, dev):
    import torch
    from subword_tokenizers_b200._lib import SwtError
    tab = P.BpeTables([("a", "b")])
    benc = dev.BpeEncoder(tab)
    arena, off = P.pack_words(["xyz"] * 1000)
    d_arena = torch.from_numpy(arena).cuda()
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).cuda()
    d_ids, d_tok, d_status = benc.encode_device(d_arena, d_off, 1000, 0, out_cap=100)
    with pytest.raises(SwtError):
        benc.check_status(d_status)


def test_encode_host_without_token_offsets(P, dev):
    import torch
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    benc = dev.BpeEncoder(tab)
    words = [w for l in load_golden("pan_tadeusz.json.gz") for w in l.lower().split()]
    arena, off = P.pack_words(words)
    ids, _, _ = benc.encode_packed(arena, off.astype(np.uint32))
    h_ids = torch.empty(len(arena) + len(words) + 16, dtype=torch.int32).pin_memory()
    nt, _ = benc.encode_host(torch.from_numpy(arena).pin_memory(), torch.from_numpy(off.astype(np.uint32).view(np.int32)).pin_memory(),
                             h_ids, None, batch_bytes=1 << 16)
    assert nt == len(ids) and np.array_equal(h_ids.numpy()[:nt].view(np.uint32), ids)


def test_encode_host_16bit_ids(P, dev):
    """swt_encode_host16: the same ids as the 32-bit call, as u16; an id that does not fit fails loudly."""
    import torch
    from subword_tokenizers_b200._lib import SwtError
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    tab = P.WpTables(load_golden("pretrained_wp_vocab.json.gz"))
    wenc = dev.WpEncoder(tab, naive_wp_encode_ids("##", tab))
    words = [w for l in load_golden("pan_tadeusz.json.gz") for w in l.lower().split()]
    arena, off = P.pack_words(words)
    ids, tok, _ = wenc.encode_packed(arena, off.astype(np.uint32))
    h_arena = torch.from_numpy(arena).pin_memory()
    h_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).pin_memory()
    h16 = torch.zeros(len(arena) + len(words) + 16, dtype=torch.int16).pin_memory()
    h_tok = torch.zeros(len(words) + 1, dtype=torch.int32).pin_memory()
    nt, _ = wenc.encode_host(h_arena, h_off, h16, h_tok, batch_bytes=1 << 16)           # many batches, odd token bases
    assert nt == len(ids)
    assert np.array_equal(h16.numpy()[:nt].view(np.uint16).astype(np.uint32), ids)
    assert np.array_equal(h_tok.numpy().view(np.uint32), tok)
    nt2, _ = wenc.encode_host(h_arena, h_off, h16, None)                                # flat token list only
    assert nt2 == nt and np.array_equal(h16.numpy()[:nt].view(np.uint16).astype(np.uint32), ids)
    # BPE: a character outside the merge alphabet comes back as 0x40000000|cp, which has no 16-bit form
    benc = dev.BpeEncoder(P.BpeTables([("a", "b")]))
    a2, o2 = P.pack_words(["ab", "a\u4e2db"])
    with pytest.raises(SwtError):
        benc.encode_host(torch.from_numpy(a2).pin_memory(), torch.from_numpy(o2.astype(np.uint32).view(np.int32)).pin_memory(),
                         torch.zeros(64, dtype=torch.int16).pin_memory(), None)


# ---- device pre-tokenization of FastWP: must equal CPython's text.lower().split() -----------------
def _device_split(dev, P, text):
    d_arena, d_off, n_words = dev.Pretokenizer.get().split_text(text)
    off = d_off.cpu().numpy().view(np.uint32)
    arena = d_arena.cpu().numpy()[: int(off[-1])].tobytes()
    assert len(off) == n_words + 1 and off[0] == 0
    return [P.decode_utf8(arena[int(off[i]):int(off[i + 1])]) for i in range(n_words)]


def test_pretok_matches_python_on_real_text(P, dev):
    lines = load_golden("pan_tadeusz.json.gz")
    for text in ("\n".join(lines), "  ".join(lines[:50]).upper(), "", " ", "\t\n", "a", " a", "a ", "Zażółć GĘŚLĄ jaźń"):
        assert _device_split(dev, P, text) == text.lower().split()


def test_pretok_unicode_corner_cases(P, dev):
    spaces = "".join(chr(c) for c in sorted(P.PY_SPACE_CPS))
    cases = [
        spaces, "x" + "y".join(spaces) + "z",                                   # every whitespace character
        "İstanbul İİ Kelvin K Ω ẞ ȺȾ \U00010400\U00010401 ǅ ǈ ῼ",      # length-changing / one-to-many lower()
        "ΟΔΥΣΣΕΥΣ ΣΟΦΟΣ Σ ΑΣ' ΑΣ'Α Σ. ΆΣ́ 'Σ ΑΣ­Σ ΑΣ: ΑΣ:Α 1Σ ΣΣΣ ʰΣ ΑʰΣ",                 # final sigma contexts
        "a\ud800b \udfff",                                                      # lone surrogates pass through
        "éÉ" * 3000 + " " + "　".join("Abİ" for _ in range(2000)),     # multi-byte characters across tile boundaries
        "x" * 4095 + "É" + "y" * 4093 + " " + "Z" * 9000,             # characters straddling 4 KiB tiles
    ]
    for text in cases:
        assert _device_split(dev, P, text) == text.lower().split(), repr(text[:40])


def test_pretok_every_code_point(P, dev):
    cps = [c for c in range(0x110000) if not 0xD800 <= c < 0xE000]
    text = " ".join(map(chr, cps))
    assert _device_split(dev, P, text) == text.lower().split()
    # every BMP code point (and the cased supplementary blocks) on both sides of a capital sigma
    probe = [c for c in cps if c < 0x10000 or 0x10400 <= c < 0x10600 or 0x1D400 <= c < 0x1D800 or 0x1E900 <= c < 0x1E960
             or 0xE0000 <= c < 0xE0200]
    text = " ".join("a%sΣ%sb Σ%s" % (chr(c), chr(c), chr(c)) for c in probe)
    assert _device_split(dev, P, text) == text.lower().split()


def test_pretok_random_texts(P, dev):
    alphabet = [chr(c) for c in (list(range(0x20, 0x7F)) + [9, 10, 13, 0x85, 0xA0, 0x130, 0x131, 0x3A3, 0x3C3, 0x3C2, 0x391, 0x3B1, 0x27,
                                                          0x301, 0xAD, 0x1E9E, 0xDF, 0x212A, 0x2003, 0x3000, 0x4E2D, 0x10400, 0x1F600,
                                                          0x141, 0x142, 0x17B, 0x41, 0x5A, 0x2C65, 0x23A])]
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 5, 127, 128, 129, 4096, 50_000):
        for _ in range(4):
            text = "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))
            assert _device_split(dev, P, text) == text.lower().split(), repr(text[:60])


def test_fastwp_tokenize_text_on_device_equals_word_path(P, dev):
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    tab = P.WpTables(load_golden("pretrained_wp_vocab.json.gz"))
    wenc = dev.WpEncoder(tab, naive_wp_encode_ids("##", tab))
    text = "\n".join(load_golden("pan_tadeusz.json.gz")).upper()
    ids_w, tok_w, _ = wenc.encode_words(text.lower().split())
    ids_t, tok_t = wenc.encode_text(text, return_offsets=True)
    assert np.array_equal(ids_w, ids_t) and np.array_equal(tok_w, tok_t)


def test_wp_tokenize_host_from_raw_text(P, dev):
    """swt_tokenize_text_host: raw text in a host buffer, many small batches, 16- and 32-bit ids."""
    import torch
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    tab = P.WpTables(load_golden("pretrained_wp_vocab.json.gz"))
    wenc = dev.WpEncoder(tab, naive_wp_encode_ids("##", tab))
    text = "\n".join(load_golden("pan_tadeusz.json.gz")) + " ΟΔΥΣΣΕΥΣ İstanbul　x"
    ids, _, _ = wenc.encode_words(text.lower().split())
    data = np.frombuffer(P.encode_utf8(text), dtype=np.uint8)
    h_text = torch.from_numpy(data.copy()).pin_memory()
    for dtype in (torch.int16, torch.int32):
        h_ids = torch.zeros(len(data) * 2 + 64, dtype=dtype).pin_memory()
        nt, nw, _ = wenc.tokenize_host(h_text, len(data), h_ids, batch_bytes=1 << 16)
        assert nw == len(text.lower().split()) and nt == len(ids)
        got = h_ids.numpy()[:nt]
        got = got.view(np.uint16).astype(np.uint32) if dtype == torch.int16 else got.view(np.uint32)
        assert np.array_equal(got, ids)


# ---- device BERT pre-tokenization (BPE classes): must equal tokenizers' BertPreTokenizer on text.lower() ------------------
def _device_bert_split(dev, P, text):
    from subword_tokenizers_b200 import _lib
    d_arena, d_off, n_words = dev.Pretokenizer.get(mode=_lib.PRETOK_BERT).split_text(text)
    off = d_off.cpu().numpy().view(np.uint32)
    arena = d_arena.cpu().numpy()[: int(off[-1])].tobytes()
    return [P.decode_utf8(arena[int(off[i]):int(off[i + 1])]) for i in range(n_words)]


def _bert_words(text):
    from tokenizers import pre_tokenizers
    return [w for w, _ in pre_tokenizers.BertPreTokenizer().pre_tokenize_str(text.lower())]


def test_bert_pretok_matches_tokenizers_library(P, dev):
    from tokenizers import pre_tokenizers
    assert P.bert_pretokenizer_matches(pre_tokenizers.BertPreTokenizer())
    lines = load_golden("pan_tadeusz.json.gz")
    cases = ["\n".join(lines), "  ".join(lines[:50]).upper(), "", " ", "a", "!", "!!", "a!b", " a,b.c ", "«Zażółć» GĘŚLĄ – jaźń…", "x\x1cy\x1fz",
             "İstanbul,İİ.Kelvin K;Ω ẞ ȺȾ", "ΟΔΥΣΣΕΥΣ, ΣΟΦΟΣ. Σ ΑΣ' ΑΣ'Α «Σ» ΑΣ:Α", "日本語。テスト、です！ 「x」", "¿qué? ¡sí!",
             "é" * 4095 + "!" + "É" * 4097 + "…" + "x" * 5000 + "、" + "y"]
    for text in cases:
        assert _device_bert_split(dev, P, text) == _bert_words(text), repr(text[:40])


def test_bert_pretok_every_code_point(P, dev):
    cps = [c for c in range(0x110000) if not 0xD800 <= c < 0xE000]
    text = "".join("a%sb%s " % (chr(c), chr(c)) for c in cps)
    assert _device_bert_split(dev, P, text) == _bert_words(text)


def test_bert_pretok_random_texts(P, dev):
    alphabet = [chr(c) for c in (list(range(0x20, 0x7F)) + [9, 10, 13, 0x1C, 0x85, 0xA0, 0xA1, 0xAB, 0xBB, 0x130, 0x3A3, 0x3C3, 0x391, 0x27, 0x301, 0x37E,
                                                          0x2003, 0x2013, 0x2014, 0x2026, 0x3000, 0x3001, 0x4E2D, 0x10400, 0x1F600, 0x141, 0x142, 0x17B,
                                                          0x1DA87, 0xFF01, 0xFE50])]
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 5, 127, 128, 129, 4096, 50_000):
        for _ in range(4):
            text = "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))
            assert _device_bert_split(dev, P, text) == _bert_words(text), repr(text[:60])


def test_fastbpe_tokenize_text_on_device_equals_word_path(P, dev):
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    benc = dev.BpeEncoder(tab)
    text = "\n".join(load_golden("pan_tadeusz.json.gz")).upper() + " " + "x" * 100 + "-" + "ab" * 40
    ids_w, tok_w, _ = benc.encode_words(_bert_words(text))
    ids_t, tok_t = benc.encode_text(text, return_offsets=True)
    assert np.array_equal(ids_w, ids_t) and np.array_equal(tok_w, tok_t)


def test_bpe_tokenize_text_host_and_small_text_path(P, dev):
    """swt_tokenize_text_host with the BPE table (BERT pre-tokenization), and the single-call path of encode_text."""
    import torch
    tab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    benc = dev.BpeEncoder(tab)
    text = "\n".join(load_golden("pan_tadeusz.json.gz")) + " ΟΔΥΣΣΕΥΣ, İstanbul!　x " + "q" * 70 + "."
    ids, _, _ = benc.encode_words(_bert_words(text))
    data = np.frombuffer(P.encode_utf8(text), dtype=np.uint8)
    h_ids = torch.zeros(len(data) * 2 + 64, dtype=torch.int32).pin_memory()
    nt, nw, _ = benc.tokenize_host(torch.from_numpy(data.copy()).pin_memory(), len(data), h_ids, batch_bytes=1 << 16)
    assert nw == len(_bert_words(text)) and nt == len(ids) and np.array_equal(h_ids.numpy()[:nt].view(np.uint32), ids)
    assert np.array_equal(benc.encode_text(text), ids)                       # <= 1 MiB: one C call
    assert np.array_equal(benc._encode_text_resident(text), ids)
    big = (text + " ") * 30                                                  # > 1 MiB: resident path
    assert len(P.encode_utf8(big)) > benc.SMALL_TEXT_BYTES
    assert np.array_equal(benc.encode_text(big), np.tile(ids, 30))


# ---- the Naive encoders as kernels ---------------------------------------------------------------------
def test_naive_encoders_on_device_match_golden_tokens(P, dev):
    """NaiveBPE.tokenize / NaiveWP.tokenize through the device path reproduce data/pan_tadeusz.tokens.json."""
    from subword_tokenizers_b200 import NaiveBPE, NaiveWP, make_hf_tokenizer
    lines = load_golden("pan_tadeusz.json.gz")
    gold = load_golden("pan_tadeusz.tokens.json.gz")
    hf = make_hf_tokenizer()
    nb = NaiveBPE(hf); nb.merges_list = [tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")]
    nw = NaiveWP(hf); nw.vocab = set(load_golden("pretrained_wp_vocab.json.gz"))
    assert nb._naive_device_encoder() is not None and nb._device_pretok_ok()
    for k in range(0, len(lines), 7):
        assert nb.tokenize(lines[k]) == gold["NaiveBPE"][k], k
        assert nw.tokenize(lines[k]) == gold["NaiveWordPiece"][k], k
    text = "\n".join(lines)
    assert nb.tokenize(text) == [t for l in gold["NaiveBPE"] for t in l]
    assert nw.tokenize(text) == [t for l in gold["NaiveWordPiece"] for t in l]


def test_naive_encoders_on_device_match_reference_on_random_cases(P, dev):
    """180 randomized merge lists / vocabularies (tests/golden/ref_random_cases.json.gz, outputs of the unmodified reference):
    untrainable merge orders (Naive != Fast), tiny alphabets, '#'-heavy vocabularies."""
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    cases = load_golden("ref_random_cases.json.gz")
    n_bpe = n_wp = n_diff = 0
    for case in cases["bpe_encode"]:
        pairs = [tuple(p) for p in case["merges"]]
        if len(set(pairs)) != len(pairs):
            continue                                   # repeated pair: the class keeps the host replay
        tab = P.BpeTables(pairs)
        arena, off = P.pack_words(case["words"])
        ids, tok, _ = dev.BpeEncoder(tab, naive=True).encode_packed(arena, off.astype(np.uint32))
        strs = tab.tokens_to_strs(ids)
        got = [strs[int(tok[i]):int(tok[i + 1])] for i in range(len(case["words"]))]
        assert got == case["naive"]
        n_bpe += 1
        n_diff += case["naive"] != case["fast"]
    for case in cases["wp_encode"]:
        tab = P.WpTables(case["vocab"])
        enc = dev.WpEncoder(tab, naive_wp_encode_ids("##", tab), naive=True)
        for text, naive in zip(case["texts"], case["naive"]):
            if naive is None:
                continue                               # the reference does not terminate on this input
            ids = enc._encode_text_resident(text)
            assert tab.tokens_to_strs(ids) == naive, (case["vocab"], text)
            n_wp += 1
    assert n_bpe >= 20 and n_diff >= 1 and n_wp >= 500


def test_tokenize_batch_equals_per_text_calls_for_all_four_classes(P, dev):
    """The harness batch entry: one pass over the concatenated texts, cut per text on the device-reported
    word positions; must equal tokenize() text by text and the reference's golden token lists."""
    from subword_tokenizers_b200 import NaiveBPE, FastBPE, NaiveWP, FastWP, make_hf_tokenizer
    from subword_tokenizers_b200.utils import WPTrie_E2E
    lines = load_golden("pan_tadeusz.json.gz")
    gold = load_golden("pan_tadeusz.tokens.json.gz")
    hf = make_hf_tokenizer()
    merges = [tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")]
    vocab = set(load_golden("pretrained_wp_vocab.json.gz"))
    nb = NaiveBPE(hf); nb.merges_list = merges
    fb = FastBPE(hf); fb.merges_list = merges; fb._rebuild_ranks()
    nw = NaiveWP(hf); nw.vocab = vocab
    fw = FastWP(hf); fw.vocab = vocab; fw.vocab_trie = WPTrie_E2E(vocab)
    texts = lines[:200] + ["", "   ", "ΟΔΥΣΣΕΥΣ", "x"]
    for tok, key in ((nb, "NaiveBPE"), (fb, "FastBPE"), (nw, "NaiveWordPiece"), (fw, "FastWordPiece")):
        got = tok.tokenize_batch(texts)
        assert got[:200] == gold[key][:200], key
        assert got[200:] == [tok.tokenize(t) for t in texts[200:]], key
        assert got[200] == [] and got[201] == []
    # texts that exceed the per-pass byte budget together are processed in several passes with the same result
    enc = fw.vocab_trie.encoder
    ids_1, cut_1 = enc.encode_texts(texts)
    enc.BATCH_TEXT_BYTES = 3000
    ids_n, cut_n = enc.encode_texts(texts)
    del enc.BATCH_TEXT_BYTES
    assert np.array_equal(ids_1, ids_n) and np.array_equal(cut_1, cut_n)


# ---- word-type table built on the device (pre-tokenization + dedupe in front of the trainers) ------------------------------
def test_device_train_types_equal_host_types(P, dev):
    from tokenizers import pre_tokenizers
    pre = pre_tokenizers.BertPreTokenizer()
    corpora = [load_golden("train-5K.json.gz"), ["a b a", "B, a! ΣΑΣ ςσ", "", "İ i̇"], ["x"], [" "], []]
    rng = np.random.default_rng(3)
    alphabet = list("abcAB .,ąŁ") + ["\n", "中", "\U00010400"]
    corpora += [["".join(alphabet[i] for i in rng.integers(0, len(alphabet), 200)) for _ in range(50)]]
    for corpus in corpora:
        words = [w for ex in corpus for w, _ in pre.pre_tokenize_str(ex.lower())]
        host = P.TrainTypes(words)
        got = dev.device_train_types(corpus, wordpiece=False)
        assert got.alphabet == host.alphabet
        assert np.array_equal(got.freq, host.freq) and np.array_equal(got.off, host.off) and np.array_equal(got.syms, host.syms)
        assert got.types == host.types
        hw = P.WpTrainTypes(words)
        gw = dev.device_train_types(corpus, wordpiece=True)
        assert sorted(gw.init_syms) == sorted(hw.init_syms)
        assert np.array_equal(gw.freq, hw.freq) and np.array_equal(gw.off, hw.off)
        assert [gw.init_syms[i] for i in gw.syms] == [hw.init_syms[i] for i in hw.syms]


def test_train_from_corpus_on_device_matches_reference(P, dev):
    """train(corpus) with pre-tokenization, type counting and the merge loop all on the GPU: the reference's merge lists /
    vocabularies on its randomized corpora and on train-5K at max_vocab 1000."""
    from subword_tokenizers_b200 import NaiveBPE, NaiveWP, make_hf_tokenizer
    hf = make_hf_tokenizer()
    cases = load_golden("ref_random_cases.json.gz")
    nb, nw = NaiveBPE(hf), NaiveWP(hf)
    assert nb._device_pretok_ok()
    for case in cases["bpe_train"]:
        nb.train(case["corpus"], case["max_vocab"])
        assert [list(m) for m in nb.merges_list] == [list(m) for m in case["merges"]]
        assert len(nb.vocab) == case["vocab_size"]
    for case in cases["wp_train"]:
        nw.train(case["corpus"], case["max_vocab"])
        assert nw.vocab == set(case["vocab"])
    corpus = load_golden("train-5K.json.gz")
    nb.train(corpus, 1000)
    assert [list(m) for m in nb.merges_list] == load_golden("ref_bpe_train5k_v1000_merges.json.gz")
    nw.train(corpus, 1000)
    assert nw.vocab == set(load_golden("ref_wp_train5k_v1000_vocab.json.gz"))
