"""Adversarial / heavy-tail inputs of the config-5 test (BASELINE configs[4]; SURVEY.md §8d "Adversarial (C5)"): deterministic, shared by
the fixture generator (unmodified reference, build container) and the GPU test (drop-in classes)."""
import numpy as np


def config5_sentences(types, vocab):
    """types: the synthetic word types the models were trained on (list of str); vocab: the trained WordPiece vocabulary.
    -> list of sentences.  Punctuation is restricted to characters that are vocabulary entries (the reference's FastWP does not
    terminate on a boundary character that is not a child of the trie root)."""
    rng = np.random.Generator(np.random.PCG64(55))
    punct = [c for c in ".,-!?'" if c in vocab]
    letters = [c for c in "abcdefghijklmnoprstuwyz" if c in vocab]
    out = []
    # words of 1, 2, 31-33, 255-257, 4 K characters; runs of one repeated character (H2); mixed with punctuation
    for n in (1, 2, 31, 32, 33, 255, 256, 257, 4096):
        out.append(" ".join(["".join(rng.choice(letters, size=n)), letters[0] * n, (letters[0] + letters[1]) * (n // 2 + 1)]))
    out.append("".join(rng.choice(letters, size=16384)))                       # one very long word (the 64 K case is in the oracle tests)
    out.append(" ".join((p.join(rng.choice(letters, size=3)) for p in punct)) if punct else "a b c")
    # heavy tail: Zipf with s = 0.6 over the word types, 40 sentences of 25 words
    n_types = len(types)
    w = 1.0 / np.arange(1, n_types + 1) ** 0.6
    cdf = np.cumsum(w); cdf /= cdf[-1]
    draw = np.searchsorted(cdf, rng.random(1000))
    words = [types[i] for i in draw]
    out += [" ".join(words[k:k + 25]) for k in range(0, 1000, 25)]
    # characters outside the training alphabet inside words (BPE: single-character tokens; WP: ['UNK'] for the whole chunk)
    out.append(" ".join(types[i] + "ξ" + types[i + 1] for i in range(0, 40, 2)))
    out.append(" ".join("ж" + types[i] for i in range(40, 60)))
    # upper case and mixed case (both tokenizers lower-case)
    out.append(" ".join(t.upper() if k % 2 else t.capitalize() for k, t in enumerate(types[:50])))
    return out
