"""Latency of small tokenize() calls (BASELINE config 1: pan_tadeusz, 989 lines) vs one tokenize_batch() call."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from subword_tokenizers_b200 import FastBPE, FastWP, make_hf_tokenizer

lines = bench.load_golden("pan_tadeusz.json.gz")
hf = make_hf_tokenizer()
bpe = FastBPE(hf); bpe.merges_list = [tuple(p) for p in bench.load_golden("pretrained_bpe_merges.json.gz")]; bpe._rebuild_ranks()
wp = FastWP(hf); wp.vocab = set(bench.load_golden("pretrained_wp_vocab.json.gz"))
from subword_tokenizers_b200.utils import WPTrie_E2E
wp.vocab_trie = WPTrie_E2E(wp.vocab)
out = {}
for name, tok in (("FastBPE", bpe), ("FastWP", wp)):
    tok.tokenize(lines[0]); tok.tokenize_batch(lines[:4])
    torch.cuda.synchronize()
    t = time.perf_counter(); per = [tok.tokenize(l) for l in lines]; t1 = time.perf_counter() - t
    t = time.perf_counter(); bat = tok.tokenize_batch(lines); t2 = time.perf_counter() - t
    t = time.perf_counter(); whole = tok.tokenize("\n".join(lines)); t3 = time.perf_counter() - t
    assert per == bat and [x for l in per for x in l] == whole
    out[name] = {"per_line_calls_s": t1, "us_per_call": 1e6 * t1 / len(lines), "one_batch_call_s": t2, "one_text_call_s": t3, "tokens": len(whole)}
print(json.dumps(out))
