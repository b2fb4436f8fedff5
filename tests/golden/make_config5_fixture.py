#!/usr/bin/env python
"""Fixture of the config-5 test: the UNMODIFIED reference (build container, /root/reference) loads the two 50 K models that
make_config5_models.py trained on the GPU box, tokenizes the adversarial inputs with all four of its classes and runs its own
token_sequence_equivalence (source/benchmarks.py:113-184) on the three model pairs the CLI compares.  Writes config5_fixture.json.gz:
the eight-tuples of the reference plus a sha256 of every class's token lists.

    python tests/golden/make_config5_fixture.py <dir with config5_bpe_merges.json.gz, config5_wp_vocab.json.gz>
"""
import gzip, hashlib, json, os, shutil, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import refshim
from config5_inputs import config5_sentences
import bench_data as BD

src = sys.argv[1] if len(sys.argv) > 1 else HERE
for name in ("config5_bpe_merges.json.gz", "config5_wp_vocab.json.gz"):
    if os.path.abspath(src) != HERE:
        shutil.copy(os.path.join(src, name), os.path.join(HERE, name))
merges = json.load(gzip.open(os.path.join(HERE, "config5_bpe_merges.json.gz"), "rt", encoding="utf-8"))
vocab = json.load(gzip.open(os.path.join(HERE, "config5_wp_vocab.json.gz"), "rt", encoding="utf-8"))
mat, lens = BD.synth_type_table(300_000, 9)
arena, off = BD.table_to_utf8(mat, lens)
types = [arena[int(off[k]):int(off[k + 1])].tobytes().decode() for k in range(2000)]
sentences = config5_sentences(types, set(vocab))

NaiveBPE, FastBPE, NaiveWP, FastWP = refshim.load_reference()
sys.path.insert(0, refshim.REFERENCE_ROOT)
from source.benchmarks import token_sequence_equivalence  # type: ignore
hf = refshim.make_hf_tokenizer()
tmp = tempfile.mkdtemp()
json.dump(merges, open(os.path.join(tmp, "merges.json"), "w", encoding="utf-8"), ensure_ascii=False)
json.dump(vocab, open(os.path.join(tmp, "vocab.json"), "w", encoding="utf-8"), ensure_ascii=False)
toks = {}
for cls in (NaiveBPE, FastBPE, NaiveWP, FastWP):
    t = cls(hf); t.load_resources(tmp); toks[cls.__name__] = t
out = {"n_sentences": len(sentences), "equivalence": {}, "token_sha256": {}, "n_tokens": {}}
for a, b in (("NaiveBPE", "FastBPE"), ("NaiveWP", "FastWP"), ("FastBPE", "FastWP")):
    out["equivalence"]["%s/%s" % (a, b)] = list(token_sequence_equivalence(toks[a], toks[b], sentences))
    print(a, b, out["equivalence"]["%s/%s" % (a, b)], flush=True)
for name, t in toks.items():
    lists = [t.tokenize(s) for s in sentences]
    out["token_sha256"][name] = hashlib.sha256(json.dumps(lists, ensure_ascii=False).encode()).hexdigest()
    out["n_tokens"][name] = sum(len(x) for x in lists)
with gzip.open(os.path.join(HERE, "config5_fixture.json.gz"), "wt", encoding="utf-8") as f:
    json.dump(out, f)
print(json.dumps(out["n_tokens"]))
