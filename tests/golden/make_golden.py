#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED upstream reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py            # minutes: copies + randomized differential cases
    python tests/golden/make_golden.py --long     # + the long trainings (BPE 1000/8000, WP 1000/8000 on
                                                  #   train-5K: ~4 min / ~30 min / ~5 min / ~45 min of CPU)

Everything written here is DATA (inputs and the reference's outputs); no reference source is
copied.  Files:

  pan_tadeusz.json.gz               reference data/pan_tadeusz.json (input lines)
  pan_tadeusz.tokens.json.gz        reference data/pan_tadeusz.tokens.json (golden, 4 models)
  train-5K.json.gz                  reference data/train-5K.json (training corpus, config 1)
  pretrained_bpe_merges.json.gz     reference resources/pretrained/FastBPE/merges.json
  pretrained_wp_vocab.json.gz       reference resources/pretrained/FastWordPiece/vocab.json
  kat_tests_resources.json          reference resources/tests/* (README tutorial KATs)
  ref_random_cases.json.gz          randomized differential cases, outputs produced by the reference
  ref_bpe_train5k_v{1000,8000}_merges.json.gz / ref_wp_train5k_v{1000,8000}_vocab.json.gz  (--long)
"""
import argparse
import gzip
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

REF = refshim.REFERENCE_ROOT


def dump_gz(name, obj):
    with gzip.open(os.path.join(HERE, name), "wt", encoding="utf-8") as f:
        json.dump(obj, f, ensure_ascii=False)
    print("wrote", name)


def load(path):
    with open(os.path.join(REF, path), encoding="utf-8") as f:
        return json.load(f)


class _Timeout(Exception):
    pass


def guarded(fn, *a, seconds=0.3):
    """Run a reference call; None when it does not terminate (SURVEY.md §7 H6 and the NaiveWP
    "#"-in-vocab remainder growth are genuine non-termination cases of the reference)."""
    import signal

    def handler(signum, frame):
        raise _Timeout()

    old = signal.signal(signal.SIGALRM, handler)
    signal.setitimer(signal.ITIMER_REAL, seconds)
    try:
        return fn(*a)
    except (_Timeout, MemoryError):
        return None
    finally:
        signal.setitimer(signal.ITIMER_REAL, 0)
        signal.signal(signal.SIGALRM, old)


def rand_word(rng, alphabet, lo, hi):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


def make_random_cases(hf):
    NaiveBPE, FastBPE, NaiveWP, FastWP = refshim.load_reference()
    rng = random.Random(20261018)
    cases = {"bpe_train": [], "bpe_encode": [], "wp_train": [], "wp_encode": []}

    # ---- HP-3: tiny alphabets force count ties (H1), runs (H2) and duplicate strings (H3)
    alphabets = ["ab", "abc", "aab", "abcdefgh", "aąbcćdeę", "xyz0123"]
    for k in range(60):
        alpha = rng.choice(alphabets)
        n_sent = rng.randint(1, 6)
        corpus = [" ".join(rand_word(rng, alpha, 1, rng.choice([4, 8, 16])) for _ in range(rng.randint(1, 12)))
                  for _ in range(n_sent)]
        if k % 7 == 0:
            corpus.append("a" * rng.randint(2, 40))
        max_vocab = rng.choice([5, 8, 12, 20, 40, 200])
        tok = FastBPE(hf)
        tok.train(corpus, max_vocab)
        cases["bpe_train"].append({"corpus": corpus, "max_vocab": max_vocab, "merges": tok.merges_list,
                                   "vocab_size": len(tok.vocab)})

    # ---- HP-1: trained lists and arbitrary (untrainable-order) lists (H9), long words, unknown chars
    for k in range(40):
        alpha = rng.choice(alphabets)
        if k % 2 == 0:
            corpus = [" ".join(rand_word(rng, alpha, 1, 10) for _ in range(30))]
            tr = FastBPE(hf)
            tr.train(corpus, rng.choice([10, 30, 80]))
            merges = list(tr.merges_list)
        else:
            syms = list(dict.fromkeys(alpha))
            merges = []
            for _ in range(rng.randint(1, 25)):
                a, b = rng.choice(syms), rng.choice(syms)
                merges.append((a, b))
                if len(a + b) < 12:
                    syms.append(a + b)
        words = [rand_word(rng, alpha + "q", 0 if k % 5 == 0 else 1, rng.choice([3, 8, 33, 70])) for _ in range(40)]
        words += ["a" * n for n in (1, 2, 3, 31, 32, 33, 64)]
        fast, naive = FastBPE(hf), NaiveBPE(hf)
        fast.merges_list = list(merges)
        fast._bpe_ranks = {pair: i for i, pair in enumerate(merges)}
        naive.merges_list = list(merges)
        cases["bpe_encode"].append({"merges": merges, "words": words,
                                    "fast": [fast.encode_word(w) for w in words],
                                    "naive": [naive.encode_word(w) for w in words]})

    # ---- NaiveWP.train on small corpora
    for k in range(40):
        alpha = rng.choice(alphabets)
        corpus = [" ".join(rand_word(rng, alpha, 1, rng.choice([4, 8])) for _ in range(rng.randint(1, 12)))
                  for _ in range(rng.randint(1, 5))]
        max_vocab = rng.choice([8, 12, 20, 40, 100])
        tok = NaiveWP(hf)
        tok.train(corpus, max_vocab)
        cases["wp_train"].append({"corpus": corpus, "max_vocab": max_vocab, "vocab": sorted(tok.vocab)})

    # ---- HP-2: FastWP on texts whose punctuation is always a vocab entry (the reference hangs
    #      otherwise); unknown letters and digits are fine (-> ['UNK'])
    for k in range(40):
        puncts = ".,!-'#" if k % 4 == 0 else ".,!-'"
        alpha = rng.choice(alphabets)
        vocab = set(puncts) | {"##" + p for p in puncts if k % 3 == 0}
        for _ in range(rng.randint(3, 40)):
            w = rand_word(rng, alpha, 1, 6)
            vocab.add(w if rng.random() < 0.5 else "##" + w)
        for c in alpha:
            if rng.random() < 0.7:
                vocab.add(c)
            if rng.random() < 0.7:
                vocab.add("##" + c)
        if k % 4 == 0:
            vocab.add("##")
        if k % 5 == 0:
            vocab.add("a.b")
            vocab.add("##.a")
        vocab = sorted(vocab)
        texts = []
        for _ in range(30):
            parts = []
            for _ in range(rng.randint(0, 8)):
                w = rand_word(rng, alpha + "zQ", 1, rng.choice([3, 8, 40]))
                r = rng.random()
                if r < 0.2:
                    w += rng.choice(puncts)
                elif r < 0.3:
                    w = rng.choice(puncts) + w
                elif r < 0.4:
                    w = w[: len(w) // 2] + rng.choice(puncts) + w[len(w) // 2:]
                elif r < 0.45:
                    w = "##" + w
                elif r < 0.5:
                    w = "##"
                parts.append(w)
            texts.append(rng.choice([" ", "  ", "\t", " \n"]).join(parts))
        tok = FastWP(hf)
        tok.vocab = set(vocab)
        from source.utils import WPTrie_E2E  # type: ignore
        tok.vocab_trie = WPTrie_E2E(tok.vocab)
        naive = NaiveWP(hf)
        naive.vocab = set(vocab)
        cases["wp_encode"].append({"vocab": vocab, "texts": texts,
                                   "fast": [guarded(tok.tokenize, t) for t in texts],
                                   "naive": [guarded(naive.tokenize, t) for t in texts]})
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long", action="store_true")
    args = ap.parse_args()
    hf = refshim.make_hf_tokenizer()

    dump_gz("pan_tadeusz.json.gz", load("data/pan_tadeusz.json"))
    dump_gz("pan_tadeusz.tokens.json.gz", load("data/pan_tadeusz.tokens.json"))
    dump_gz("train-5K.json.gz", load("data/train-5K.json"))
    dump_gz("pretrained_bpe_merges.json.gz", load("resources/pretrained/FastBPE/merges.json"))
    dump_gz("pretrained_wp_vocab.json.gz", load("resources/pretrained/FastWordPiece/vocab.json"))
    kat = {name: load("resources/tests/%s/%s" % (name, "merges.json" if "BPE" in name else "vocab.json"))
           for name in ("NaiveBPE", "FastBPE", "NaiveWordPiece", "FastWordPiece")}
    kat["corpus"] = ["This is a sentence.", "Another example sentence."]          # README.md:145-147
    kat["max_vocab"] = 25
    NaiveBPE, FastBPE, NaiveWP, FastWP = refshim.load_reference()
    sent = "This sentence is a new example."                                       # README.md:167-172
    out = {}
    for name, cls in (("NaiveBPE", NaiveBPE), ("FastBPE", FastBPE), ("NaiveWordPiece", NaiveWP), ("FastWordPiece", FastWP)):
        t = cls(hf)
        t.load_resources(os.path.join(REF, "resources/tests", name))
        out[name] = t.tokenize(sent)
    kat["readme_sentence"] = sent
    kat["readme_tokens"] = out
    with open(os.path.join(HERE, "kat_tests_resources.json"), "w", encoding="utf-8") as f:
        json.dump(kat, f, ensure_ascii=False, indent=1)
    print("wrote kat_tests_resources.json")

    t0 = time.time()
    dump_gz("ref_random_cases.json.gz", make_random_cases(hf))
    print("random cases: %.1f s" % (time.time() - t0))

    if args.long:
        corpus = load("data/train-5K.json")
        for vocab in (1000, 8000):
            tok = FastBPE(hf)
            tok.train(corpus, vocab)
            dump_gz("ref_bpe_train5k_v%d_merges.json.gz" % vocab, tok.merges_list)
        for vocab in (1000, 8000):
            tok = NaiveWP(hf)
            tok.train(corpus, vocab)
            dump_gz("ref_wp_train5k_v%d_vocab.json.gz" % vocab, sorted(tok.vocab))


if __name__ == "__main__":
    main()
