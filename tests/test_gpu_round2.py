"""Round-2 GPU parity tests (pytest -m gpu, run on the B200 box): the memo beyond 2^20 word types, a contention stress of the memo
protocol, long chunks split across a warp, the FastBPE config-1 hash on the CUDA path, punctuation-dense text through the host
pipeline, large alphabets and empty corpora in the trainers, live-container tracking of the classes, and the multi-GPU trainer over
NCCL / peer memory (skipped below 2 GPUs).  Everything is compared bit for bit with the CPU oracle or the reference's golden data."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from subword_tokenizers_b200 import packing
    return packing


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from subword_tokenizers_b200 import device
    return device


def _wp(P, dev, vocab):
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    tab = P.WpTables(vocab)
    return tab, dev.WpEncoder(tab, naive_wp_encode_ids("##", tab))


def _oracle_wp(P, tab, arena, off):
    import oracle
    alnum, space = P.unicode_class_bitmaps()
    return oracle.WpTrie(tab, alnum).encode(arena, off, space)


# ------------------------------------------------------------------------------------------------ memo
def test_more_than_2_20_distinct_word_types(P, dev):
    """1.5 M distinct types + a Zipf head: more types than the round-1 memo could hold (2^20); FastWP and FastBPE vs the oracle."""
    import oracle
    import bench_data as BD
    mat, lens = BD.synth_type_table(1_500_000, 7)
    t_arena, t_off = BD.table_to_utf8(mat, lens)
    rng = np.random.default_rng(3)
    n_types = len(lens)
    draw = np.concatenate([np.arange(n_types), rng.integers(0, 2000, size=1_500_000), rng.integers(0, n_types, size=500_000)])
    rng.shuffle(draw)
    tl = np.diff(t_off)
    off = np.zeros(len(draw) + 1, dtype=np.int64)
    np.cumsum(tl[draw], out=off[1:])
    idx = np.repeat(t_off[:-1][draw] - off[:-1], tl[draw]) + np.arange(off[-1])
    arena = t_arena[idx]
    tab, enc = _wp(P, dev, load_golden("pretrained_wp_vocab.json.gz"))
    ids, tok_off, h6 = enc.encode_packed(arena, off.astype(np.uint32))
    o_ids, o_off, o_h6 = _oracle_wp(P, tab, arena, off.astype(np.uint64))
    assert np.array_equal(tok_off.astype(np.uint64), o_off) and np.array_equal(ids, o_ids) and h6 == o_h6
    btab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    benc = dev.BpeEncoder(btab)
    ids, tok_off, _ = benc.encode_packed(arena, off.astype(np.uint32))
    o_ids, o_off = oracle.bpe_encode(btab, arena, off.astype(np.uint64))
    assert np.array_equal(tok_off.astype(np.uint64), o_off) and np.array_equal(ids, o_ids)
    # the same with a memo too small for the stream (2^12 slots): overflowing types take the direct path
    dev.tune("memo_max_log2", 12)
    try:
        ids, tok_off, _ = benc.encode_packed(arena[: int(off[400_000])], off[:400_001].astype(np.uint32))
        o_ids, o_off = oracle.bpe_encode(btab, arena[: int(off[400_000])], off[:400_001].astype(np.uint64))
        assert np.array_equal(tok_off.astype(np.uint64), o_off) and np.array_equal(ids, o_ids)
        ids, tok_off, h6 = enc.encode_packed(arena[: int(off[400_000])], off[:400_001].astype(np.uint32))
        o_ids, o_off, o_h6 = _oracle_wp(P, tab, arena[: int(off[400_000])], off[:400_001].astype(np.uint64))
        assert np.array_equal(tok_off.astype(np.uint64), o_off) and np.array_equal(ids, o_ids)
    finally:
        dev.tune("memo_max_log2", 22)


def test_split_count_pass_leaf_and_resolve_kernels(P, dev):
    """FastBPE / NaiveBPE with the split count pass (warm-up kernel, then leaf count + resolve kernels per chunk) forced onto a small
    stream: many first occurrences after the warm-up, words of 16..32 and more than 32 bytes, empty words, a partial last tile."""
    import oracle
    import bench_data as BD
    mat, lens = BD.synth_type_table(120_000, 3)
    t_arena, t_off = BD.table_to_utf8(mat, lens)
    rng = np.random.default_rng(29)
    draw = np.concatenate([rng.integers(0, 500, size=150_000), np.arange(120_000), rng.integers(0, 120_000, size=100_000)])
    rng.shuffle(draw)
    # a prefix of few types: the warm-up (64 tiles here) sees almost no slow-path words, so the split form is taken; without the
    # prefix (second pass below) the warm-up meets mostly first occurrences and the device-side gate hands the stream to the single kernel
    draw = np.concatenate([rng.integers(0, 20, size=6000), draw])
    words = [t_arena[int(t_off[i]):int(t_off[i + 1])].tobytes().decode() for i in draw[:60_000]]
    words += ["", "x" * 40, "wielkopolskiego" + "ab", "a" * 33, "zażółćgęśląjaźńzażółć"] * 50
    extra_arena, extra_off = P.pack_words(words)
    tl = np.diff(t_off)
    off = np.zeros(len(draw) + 1, dtype=np.int64)
    np.cumsum(tl[draw], out=off[1:])
    idx = np.repeat(t_off[:-1][draw] - off[:-1], tl[draw]) + np.arange(off[-1])
    arena = np.concatenate([t_arena[idx], extra_arena])
    off = np.concatenate([off, off[-1] + extra_off[1:].astype(np.int64)])
    merges = [tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")]
    btab = P.BpeTables(merges)
    dev.tune("split_warm_tiles", 64); dev.tune("split_chunks", 3)
    try:
        for naive in (False, True):
            tab = btab if not naive else P.BpeTables(merges[:4000])
            enc = dev.BpeEncoder(tab, naive=naive)
            o_ids, o_off = oracle.bpe_encode(tab, arena, off.astype(np.uint64), naive=naive)
            for skip in (0, 6000):                                      # with / without the prefix of few types
                a0, t0 = int(off[skip]), int(o_off[skip])
                ids, tok_off, _ = enc.encode_packed(arena[a0:], (off[skip:] - off[skip]).astype(np.uint32))
                assert np.array_equal(tok_off.astype(np.uint64), o_off[skip:] - o_off[skip]) and np.array_equal(ids, o_ids[t0:]), (naive, skip)
            enc.close()
    finally:
        dev.tune("split_warm_tiles", 16384); dev.tune("split_chunks", 2)


def test_memo_contention_stress(P, dev):
    """The memo protocol under contention (compute-sanitizer is closed on this pool, so the check is a stress test): a handful of
    16..32-byte types that share their 15-byte key prefix, hot short types, and thousands of first occurrences, on every SM at once,
    200 runs, each compared with the oracle.  FastWP and FastBPE (both miss-resolution modes)."""
    import oracle
    import torch
    rng = np.random.default_rng(17)
    prefix = "wielkopolskiego"                                        # 15 bytes
    shared = [prefix + s for s in ("a", "ab", "abc", "x", "xyz0123456789abcd", "b" * 17, "ąę", "1", "zz", "zzz")]
    hot = ["i", "w", "na", "się", "z", "do", "nie", "to", "że", "a"]
    alphabet = list("abcdefghijklmnoprstuwyząęłóż0123456789")
    first = ["".join(rng.choice(alphabet, size=int(rng.integers(1, 31)))) for _ in range(60_000)]
    pool = shared * 3000 + hot * 4000 + first
    order = rng.permutation(len(pool))
    words = [pool[i] for i in order]
    arena, off = P.pack_words(words)
    off32 = off.astype(np.uint32)
    wtab, wenc = _wp(P, dev, load_golden("pretrained_wp_vocab.json.gz"))
    btab = P.BpeTables([tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")])
    benc = dev.BpeEncoder(btab)
    w_ids, w_off, w_h6 = _oracle_wp(P, wtab, arena, off)
    b_ids, b_off = oracle.bpe_encode(btab, arena, off)
    d = torch.device("cuda", torch.cuda.current_device())
    d_arena = torch.from_numpy(arena).to(d)
    d_off = torch.from_numpy(off32.view(np.int32)).to(d)
    n_words = len(words)
    lens = np.diff(off.astype(np.int64))
    long_bytes = int(lens[lens > 32].sum())
    dev.tune("memo_max_log2", 17)                                     # crowded table: collisions, full probe sequences
    try:
        for kind, enc, want_ids, want_off in (("wp", wenc, w_ids, w_off), ("bpe", benc, b_ids, b_off), ("bpe-in-tile", benc, b_ids, b_off)):
            dev.tune("bpe_queue", 0 if kind == "bpe-in-tile" else 1)
            for run in range(200 if kind != "bpe-in-tile" else 60):
                d_ids, d_tok, d_status = enc.encode_device(d_arena, d_off, n_words, long_bytes)
                nt, _ = enc.check_status(d_status)
                assert nt == len(want_ids), (kind, run)
                assert torch.equal(d_tok.cpu(), torch.from_numpy(want_off.astype(np.uint32).view(np.int32))), (kind, run)
                assert torch.equal(d_ids[:nt].cpu(), torch.from_numpy(want_ids.view(np.int32))), (kind, run)
    finally:
        dev.tune("memo_max_log2", 22); dev.tune("bpe_queue", 1)


# ------------------------------------------------------------------------------------------------ long chunks (north-star item 3)
def test_wp_long_chunks_split_across_the_warp(P, dev):
    """Chunks of 33 bytes .. 64 K characters with many word boundaries: segments walked by different lanes.  Vocabularies with entries
    that cross a boundary ("a.b", "u.s.") force the sequential chain; unknown punctuation gives H6 events inside long chunks."""
    rng = np.random.default_rng(23)
    vocab = set(load_golden("pretrained_wp_vocab.json.gz"))
    for extra in ((), ("a.b", "u.s.", "##.x", "-a", "a-", "...", ".,", "a.", "##a.b")):
        tab, enc = _wp(P, dev, vocab | set(extra))
        alphabet = list("abcdeiknorstwyzłó") + list(".,-!?'") + ["€", "§"]
        chunks = []
        for n in (33, 34, 40, 64, 65, 100, 255, 256, 257, 1000, 4096, 20000, 65536):
            chunks.append("".join(rng.choice(alphabet, size=n)))
            chunks.append(".".join("ab" for _ in range(n // 3)))
            chunks.append("a.b" * (n // 3))
            chunks.append("u.s." * (n // 4) + "x")
            chunks.append("€" * (n // 3))                               # H6 at every character
            chunks.append("a" * n)                                      # one segment: sequential
            chunks.append("a-" * (n // 2))
        chunks += ["".join(rng.choice(alphabet, size=int(rng.integers(33, 90)))) for _ in range(3000)]
        chunks += ["ab", "", "x" * 32]
        ids, tok_off, h6 = enc.encode_words(chunks)
        arena, off = P.pack_words(chunks)
        o_ids, o_off, o_h6 = _oracle_wp(P, tab, arena, off)
        assert np.array_equal(tok_off.astype(np.uint64), o_off), extra
        assert np.array_equal(ids, o_ids) and h6 == o_h6 and h6 > 0, extra
        # without scratch (long_word_bytes = 0) the owning lane walks the chunk alone: same result
        import torch
        d = torch.device("cuda", torch.cuda.current_device())
        d_ids, d_tok, d_status = enc.encode_device(torch.from_numpy(arena).to(d), torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(d),
                                                   len(chunks), 0)
        nt, h6b = enc.check_status(d_status)
        assert nt == len(o_ids) and h6b == o_h6 and np.array_equal(d_ids[:nt].cpu().numpy().view(np.uint32), o_ids)
        enc.close()


# ------------------------------------------------------------------------------------------------ config 1 on the CUDA path
def test_fastbpe_config1_hash_on_the_cuda_path(hf_tokenizer):
    """BASELINE config 1: FastBPE with the 922 merges trained on train-5K (max_vocab 1000) tokenizing pan_tadeusz line by line;
    the sha256 of the token lists is the one the survey took from the unmodified reference."""
    from subword_tokenizers_b200 import FastBPE
    tok = FastBPE(hf_tokenizer)
    tok.merges_list = [tuple(p) for p in load_golden("ref_bpe_train5k_v1000_merges.json.gz")]
    tok._rebuild_ranks()
    lines = load_golden("pan_tadeusz.json.gz")
    out = [tok.tokenize(line) for line in lines]
    assert sum(len(t) for t in out) == 16469
    h = hashlib.sha256(json.dumps(out, ensure_ascii=False).encode()).hexdigest()
    assert h == "59c4833e7556ced4d1786716d3dc8fa31ac45a4c8d116fa4e159c4ca6a33ae58"
    assert tok.tokenize_batch(lines) == out


def test_per_line_tokenize_equals_golden_for_all_four_classes(hf_tokenizer, tmp_path):
    """The call pattern of cli.py:253-264 (one tokenize() per line) with the pretrained models: every line equals
    data/pan_tadeusz.tokens.json for all four classes."""
    from subword_tokenizers_b200 import FastBPE, FastWP, NaiveBPE, NaiveWP
    golden = load_golden("pan_tadeusz.tokens.json.gz")
    lines = load_golden("pan_tadeusz.json.gz")
    merges = load_golden("pretrained_bpe_merges.json.gz")
    vocab = load_golden("pretrained_wp_vocab.json.gz")
    for d, payload, name in (("bpe", merges, "merges.json"), ("wp", vocab, "vocab.json")):
        os.makedirs(tmp_path / d, exist_ok=True)
        with open(tmp_path / d / name, "w", encoding="utf-8") as f:
            json.dump(payload, f, ensure_ascii=False)
    for cls, key, d in ((FastBPE, "FastBPE", "bpe"), (FastWP, "FastWordPiece", "wp"), (NaiveWP, "NaiveWordPiece", "wp"), (NaiveBPE, "NaiveBPE", "bpe")):
        tok = cls(hf_tokenizer)
        tok.load_resources(str(tmp_path / d))
        n = len(lines) if cls is not NaiveBPE else 120
        for k in range(n):
            assert tok.tokenize(lines[k]) == golden[key][k], (key, k)


# ------------------------------------------------------------------------------------------------ ADVICE (round 1)
def test_punctuation_dense_text_through_the_host_pipeline(hf_tokenizer, pre_tokenize):
    """BERT mode makes every punctuation character a word: 'a,a,a,...' has one word per byte.  700 KB of it must tokenize
    (round 1 sized the pipeline for one word per two bytes), also when the text is cut into several batches."""
    from subword_tokenizers_b200 import FastBPE
    tok = FastBPE(hf_tokenizer)
    tok.merges_list = [("a", "b"), ("ab", "c")]
    tok._rebuild_ranks()
    text = "a," * 350_000
    out = tok.tokenize(text)
    assert len(out) == 700_000 and out[:4] == ["a", ",", "a", ","]
    text2 = ("abc,.;" * 40_000) + " ab\x1ccd " + ("!?" * 30_000)
    enc = tok._device_encoder()
    want = enc.tables.tokens_to_strs(enc.encode_words(pre_tokenize(text2))[0])
    assert tok.tokenize(text2) == want and "abc" in want
    # a multi-batch call through the C entry point with a small pipeline: 0x1C-0x1F are not cut points in BERT mode
    data = (("ab\x1ccd\x1dxy " * 30_000) + "a,b.c").encode()
    out_ids = np.empty(2 * len(data) + 16, dtype=np.uint32)
    nt, nw, _ = enc.tokenize_host(np.frombuffer(data, dtype=np.uint8), len(data), out_ids, has_sigma=False, batch_bytes=1 << 16)
    want_words = pre_tokenize(data.decode())
    assert nw == len(want_words)
    ids, _, _ = enc.encode_words(want_words)
    assert nt == len(ids) and np.array_equal(out_ids[:nt], ids)


def test_training_with_an_alphabet_beyond_4096_symbols(P, dev):
    """CJK-sized alphabets: the initial pair counts go straight into the pair table (no dense n_alpha^2 array)."""
    import oracle
    rng = np.random.default_rng(31)
    cps = np.arange(0x4E00, 0x4E00 + 5200)
    words = ["".join(chr(c) for c in rng.choice(cps, size=int(rng.integers(1, 6)))) for _ in range(30_000)]
    words += ["".join(chr(c) for c in cps[:40])] * 3 + [chr(cps[0]) * 9, chr(cps[1]) + chr(cps[0])] * 50
    tt = P.TrainTypes(words)
    assert tt.n_alpha > 4096
    max_vocab = tt.n_alpha + 300
    max_len = int(np.diff(tt.off.astype(np.int64)).max())
    eng = dev.CudaTrainEngine(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab, tt.n_alpha, max_len, 0, 0, 1, record_cap=128)
    l, r, n, c, state = dev.run_training_loop(eng, 1, steps_per_sync=128)
    ol, orr, on, oc, ovs = oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab)
    assert len(l) == len(ol) > 100
    assert np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on) and np.array_equal(c, oc)
    assert state["vocab_size"] == ovs


def test_training_large_table_paths_ties_filter_compaction(P, dev):
    """The paths only big corpora take, forced on a small one: two-level argmax cache + listed tie keys (table_cap 2^21), the pair
    filter of the mark scan, its rebuild and the word-table compaction (maintenance every 32 steps).  Heavy ties: all types have
    frequency 1 or 2.  Merge list, chosen counts and the final segmentation equal the oracle's / a replay."""
    import oracle
    rng = np.random.default_rng(41)
    alphabet = list("abcdefghijklmnopqrstuvwxyz")
    words = ["".join(rng.choice(alphabet, size=int(rng.integers(2, 14)))) for _ in range(24_000)]
    words += words[:3000]                                            # frequency 2 for the first types
    tt = P.TrainTypes(words)
    max_len = int(np.diff(tt.off.astype(np.int64)).max())
    max_vocab = 26 + 1500
    eng = dev.CudaTrainEngine(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab, tt.n_alpha, max_len, 0, 0, 1, record_cap=64, table_cap=1 << 21)
    l, r, n, c, state = dev.run_training_loop(eng, 1, steps_per_sync=32)
    ol, orr, on, oc, ovs = oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab)
    assert len(l) == len(ol) == 1500
    assert np.array_equal(c, oc) and np.array_equal(l, ol) and np.array_equal(r, orr) and np.array_equal(n, on)
    assert state["n_tie_steps"] > 500
    merges, strs = tt.merges_to_strs(l, r, n)
    syms, lens = eng.read_corpus()
    for k in range(0, tt.n_types, 211):
        word = list(tt.types[k])
        for a, b in merges:
            out, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(word[i]); i += 1
            word = out
        s0 = int(tt.off[k])
        assert [strs[j] for j in syms[s0:s0 + int(lens[k])]] == word


def test_empty_corpus_and_live_container_tracking(hf_tokenizer):
    from subword_tokenizers_b200 import FastBPE, FastWP, NaiveBPE, NaiveWP
    for cls in (NaiveBPE, FastBPE, NaiveWP, FastWP):
        tok = cls(hf_tokenizer)
        tok.train([], 100)                                           # the reference returns normally with nothing learnt
        assert len(tok.vocab) == 0
        tok.train(["", "   "], 100)
        assert len(tok.vocab) == 0
    # in-place edits that keep the length must reach the device tables (round 1 keyed its caches on id + len)
    tok = FastBPE(hf_tokenizer)
    tok.merges_list = [("a", "b"), ("c", "d")]
    tok._rebuild_ranks()
    assert tok.tokenize("abcd") == ["ab", "##cd"]
    del tok._bpe_ranks[("c", "d")]
    tok._bpe_ranks[("b", "c")] = 1                                   # same length, same first and last key ... different model
    assert tok.tokenize("abcd") == ["ab", "##c", "##d"]
    nb = NaiveBPE(hf_tokenizer)
    nb.merges_list = [("a", "b"), ("c", "d")]
    assert nb.tokenize("abcd") == ["ab", "##cd"]
    nb.merges_list[1] = ("ab", "c")
    assert nb.tokenize("abcd") == ["abc", "##d"]
    nw = NaiveWP(hf_tokenizer)
    nw.vocab = {"ab", "##cd", "a", "##b"}
    assert nw.tokenize("abcd") == ["ab", "##cd"]
    nw.vocab.discard("##cd"); nw.vocab.add("##c")
    assert nw.tokenize("abcd") == ["[UNK]"]


# ------------------------------------------------------------------------------------------------ multi-GPU trainer on hardware
def test_multi_gpu_trainer_matches_single_gpu_and_oracle():
    """CudaTrainEngine on 2 (and 4, when present) ranks, one process per GPU, NCCL + the peer-memory exchange: the merge list of
    every rank equals the single-GPU list and the oracle's (N-GPU == 1-GPU == reference)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for world in [w for w in (2, 4, 8) if w <= n]:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29650 + world), os.path.join(ROOT, "tests", "mgpu_train_worker.py")]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
        assert "MGPU_TRAIN_OK world=%d" % world in res.stdout


# ------------------------------------------------------------------------------------------------ small calls (cli.py:253-264)
def test_small_call_path_equals_the_batch_path(hf_tokenizer, tmp_path):
    """swt_tokenize_small (two single-CTA kernels, zero-copy buffers) against the resident batch path and the golden lines: texts
    around the size limit, words longer than 32 bytes, capital sigma, unknown characters, empty and whitespace-only texts."""
    from subword_tokenizers_b200 import FastBPE, FastWP, NaiveBPE, NaiveWP
    from subword_tokenizers_b200.device import SmallCall
    merges = load_golden("pretrained_bpe_merges.json.gz")
    vocab = load_golden("pretrained_wp_vocab.json.gz")
    lines = load_golden("pan_tadeusz.json.gz")
    rng = np.random.default_rng(5)
    limit = SmallCall.get().max_bytes
    texts = ["", " ", "\n\t ", "a", "Litwo! Ojczyzno moja", "ΑΣ ΑΣΑ Σ ΟΔΥΣΣΕΎΣ", "x" * 33 + " " + "y" * 200 + " tail", "a,b.c;d" * 50, "€§ zażółć GĘŚLĄ jaźń İstanbul",
             "a" * 239, "ab " * 80, "ż" * 120, "a" * 241, "Ab,c " * 48 + "x",          # around the 240 bytes that travel as kernel parameters
             "słowo " * ((limit - 8) // 7), ("ab " * (limit // 3))[:limit], ("ab " * (limit // 3))[:limit - 1] + "Z", "q" * limit, " ".join(lines[:40]),
             "".join(rng.choice(list("abcdeiknorstwyzłó .,-!?'"), size=3000))]
    for cls, payload in ((FastWP, vocab), (FastBPE, merges), (NaiveWP, vocab), (NaiveBPE, merges[:3000])):
        tok = cls(hf_tokenizer)
        if cls in (FastWP, NaiveWP):
            tok.vocab = set(payload)
            if cls is FastWP:
                from subword_tokenizers_b200.utils import WPTrie_E2E
                tok.vocab_trie = WPTrie_E2E(tok.vocab)
            enc = tok.vocab_trie.encoder if cls is FastWP else tok._naive_device_encoder()
        else:
            tok.merges_list = [tuple(p) for p in payload]
            if cls is FastBPE:
                tok._rebuild_ranks()
                enc = tok._device_encoder()
            else:
                enc = tok._naive_device_encoder()
        for t in texts:
            data = t.encode()
            assert len(data) <= limit or cls is not None
            small = enc.encode_text(t) if len(data) <= limit else None
            big = enc._encode_text_resident(t)
            if small is not None:
                assert np.array_equal(small, big), (cls.__name__, t[:30], len(data))


# ------------------------------------------------------------------------------------------------ BASELINE configs[4]: --compare at 50 K
def _token_sequence_equivalence(tok1, tok2, sentences):
    """The metric of the reference's --compare (source/benchmarks.py:113-184), restated for the test: positional matches, unordered
    overlap and per-word shared-subword rate between two tokenizers."""
    from collections import Counter
    strip = lambda toks: [t[2:] if t.startswith("##") else t for t in toks]         # noqa: E731
    pos = n_pos = unordered = n_words = word_matches = 0
    for s in sentences:
        t1, t2 = strip(tok1.tokenize(s)), strip(tok2.tokenize(s))
        n = min(len(t1), len(t2))
        pos += sum(1 for i in range(n) if t1[i] == t2[i]); n_pos += n
        f1, f2 = Counter(t1), Counter(t2)
        unordered += sum(min(f1[t], f2[t]) for t in (f1.keys() & f2.keys()))
        words = s.split()
        n_words += len(words)
        for w in words:
            if set(strip(tok1.tokenize(w))) & set(strip(tok2.tokenize(w))):
                word_matches += 1
    return [pos, n_pos, (pos / n_pos * 100) if n_pos else 0.0, unordered, (unordered / n_pos * 100) if n_pos else 0.0,
            word_matches, n_words, (word_matches / n_words * 100) if n_words else 0.0]


def test_config5_compare_equivalence_at_50k_vocab(hf_tokenizer, tmp_path):
    """BASELINE configs[4]: the four drop-in classes with 50 K models (trained by the GPU trainers on synthetic types, oracle-checked
    prefix: tests/golden/make_config5_models.py) on adversarial long-word / heavy-tail input.  Their token lists and the three
    --compare tuples must equal what the UNMODIFIED reference produced with the same models (tests/golden/make_config5_fixture.py)."""
    import bench_data as BD
    from subword_tokenizers_b200 import FastBPE, FastWP, NaiveBPE, NaiveWP
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from config5_inputs import config5_sentences
    fixture = load_golden("config5_fixture.json.gz")
    merges, vocab = load_golden("config5_bpe_merges.json.gz"), load_golden("config5_wp_vocab.json.gz")
    assert len(vocab) == 50_000 and len(merges) > 49_000
    mat, lens = BD.synth_type_table(300_000, 9)
    arena, off = BD.table_to_utf8(mat, lens)
    types = [arena[int(off[k]):int(off[k + 1])].tobytes().decode() for k in range(2000)]
    sentences = config5_sentences(types, set(vocab))
    assert len(sentences) == fixture["n_sentences"]
    for d, payload, name in (("bpe", merges, "merges.json"), ("wp", vocab, "vocab.json")):
        os.makedirs(tmp_path / d, exist_ok=True)
        with open(tmp_path / d / name, "w", encoding="utf-8") as f:
            json.dump(payload, f, ensure_ascii=False)
    toks = {}
    for cls, d in ((NaiveBPE, "bpe"), (FastBPE, "bpe"), (NaiveWP, "wp"), (FastWP, "wp")):
        t = cls(hf_tokenizer)
        t.load_resources(str(tmp_path / d))
        toks[cls.__name__] = t
        lists = [t.tokenize(s) for s in sentences]
        assert sum(len(x) for x in lists) == fixture["n_tokens"][cls.__name__], cls.__name__
        assert hashlib.sha256(json.dumps(lists, ensure_ascii=False).encode()).hexdigest() == fixture["token_sha256"][cls.__name__], cls.__name__
        assert t.tokenize_batch(sentences) == lists, cls.__name__
    for pair, want in fixture["equivalence"].items():
        a, b = pair.split("/")
        got = _token_sequence_equivalence(toks[a], toks[b], sentences)
        assert got == want, (pair, got, want)


def test_direct_path_without_memo_equals_oracle(P, dev):
    """swt_tune("memo_off", 1): no word-type memo, every word is encoded by its own lane in the count pass and again in the emit pass
    (the direct-path rates of bench.py).  All four encoders against the oracle, with empty, 16..32-byte and long words."""
    import oracle
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    rng = np.random.default_rng(37)
    alphabet = list("abcdeiknorstwyzłóżę") + list(".,-'")
    words = ["".join(rng.choice(alphabet, size=int(rng.integers(1, 24)))) for _ in range(40_000)]
    words += ["", "x" * 33, "a" * 200, "wielkopolskiegoab", "zażółćgęśląjaźń"] * 20
    arena, off = P.pack_words(words)
    merges = [tuple(p) for p in load_golden("pretrained_bpe_merges.json.gz")]
    vocab = load_golden("pretrained_wp_vocab.json.gz")
    alnum, space = P.unicode_class_bitmaps()
    dev.tune("memo_off", 1)
    try:
        for naive in (False, True):
            btab = P.BpeTables(merges if not naive else merges[:3000])
            benc = dev.BpeEncoder(btab, naive=naive)
            ids, tok_off, _ = benc.encode_packed(arena, off.astype(np.uint32))
            o_ids, o_off = oracle.bpe_encode(btab, arena, off, naive=naive)
            assert np.array_equal(tok_off.astype(np.uint64), o_off) and np.array_equal(ids, o_ids), ("bpe", naive)
            benc.close()
        wtab = P.WpTables(vocab)
        wenc = dev.WpEncoder(wtab, naive_wp_encode_ids("##", wtab))
        ids, tok_off, h6 = wenc.encode_packed(arena, off.astype(np.uint32))
        o_ids, o_off, o_h6 = oracle.WpTrie(wtab, alnum).encode(arena, off, space)
        assert np.array_equal(tok_off.astype(np.uint64), o_off) and np.array_equal(ids, o_ids) and h6 == o_h6
        wenc.close()
    finally:
        dev.tune("memo_off", 0)
