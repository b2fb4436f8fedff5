"""BASELINE configs[4] ("--compare equivalence at 50K vocab"): trains the two 50 K models of the config-5 test with the GPU trainers.

Runs on the GPU box (gpurun): a synthetic corpus (Zipf draws over bench_data.synth_type_table word types, 20 words per sentence) goes
through FastBPE.train / FastWP.train of the drop-in classes (device pre-tokenization, type counting and merge loop); the first merges of
both trainings are compared with the CPU oracle on host-built word types.  Writes config5_bpe_merges.json.gz / config5_wp_vocab.json.gz;
tests/golden/make_config5_fixture.py (build container, unmodified reference) turns them into the fixture of the test.

    python tests/golden/make_config5_models.py <out_dir>
"""
import gzip, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench_data as BD
import oracle
from subword_tokenizers_b200 import FastBPE, FastWP, make_hf_tokenizer, packing as P

out_dir = sys.argv[1] if len(sys.argv) > 1 else "."
N_TYPES, N_WORDS, VOCAB = 300_000, 3_000_000, 50_000
mat, lens = BD.synth_type_table(N_TYPES, 9)
arena, off = BD.table_to_utf8(mat, lens)
types = [arena[int(off[k]):int(off[k + 1])].tobytes().decode() for k in range(N_TYPES)]
rng = np.random.Generator(np.random.PCG64(9))
w = 1.0 / np.arange(1, N_TYPES + 1); cdf = np.cumsum(w); cdf /= cdf[-1]
draw = np.searchsorted(cdf, rng.random(N_WORDS))
words = [types[i] for i in draw]
corpus = [" ".join(words[k:k + 20]) for k in range(0, N_WORDS, 20)]
hf = make_hf_tokenizer()
report = {"n_types": N_TYPES, "n_words": N_WORDS, "max_vocab": VOCAB}

bpe = FastBPE(hf)
bpe.train(corpus, VOCAB)
tt = P.TrainTypes(words)
K = 6
ol, orr, on, oc, _ = oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, tt.n_alpha + K)
want, _ = tt.merges_to_strs(ol, orr, on)
assert [tuple(m) for m in bpe.merges_list[:len(want)]] == [tuple(m) for m in want], (bpe.merges_list[:K], want)
report["bpe_merges"] = len(bpe.merges_list); report["bpe_vocab"] = len(bpe.vocab); report["bpe_oracle_prefix"] = len(want)

wp = FastWP(hf)
wp.train(corpus, VOCAB)
wt = P.WpTrainTypes(words)
ol, orr, on, _ = oracle.wp_train(wt.syms, wt.off, wt.freq, wt.init_cps, wt.init_off, len(wt.init_syms) + K)
want_vocab = set(wt.vocab_from_merges(ol, orr, on))
assert want_vocab <= set(wp.vocab), sorted(want_vocab - set(wp.vocab))[:5]
report["wp_vocab"] = len(wp.vocab); report["wp_oracle_prefix_tokens"] = len(want_vocab) - len(wt.init_syms)

with gzip.open(os.path.join(out_dir, "config5_bpe_merges.json.gz"), "wt", encoding="utf-8") as f:
    json.dump([list(m) for m in bpe.merges_list], f, ensure_ascii=False)
with gzip.open(os.path.join(out_dir, "config5_wp_vocab.json.gz"), "wt", encoding="utf-8") as f:
    json.dump(sorted(wp.vocab), f, ensure_ascii=False)
print(json.dumps(report))
