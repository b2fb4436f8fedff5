"""Where a per-line tokenize() call spends its time: the bound C call alone, + the id copy, + the id -> string table, the class call."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from subword_tokenizers_b200 import FastWP, FastBPE, make_hf_tokenizer, device
from subword_tokenizers_b200.utils import WPTrie_E2E

lines = bench.load_golden("pan_tadeusz.json.gz")
hf = make_hf_tokenizer()
wp = FastWP(hf); wp.vocab = set(bench.load_golden("pretrained_wp_vocab.json.gz")); wp.vocab_trie = WPTrie_E2E(wp.vocab)
bpe = FastBPE(hf); bpe.merges_list = [tuple(p) for p in bench.load_golden("pretrained_bpe_merges.json.gz")]; bpe._rebuild_ranks()
out = {}
for name, tok, enc in (("FastWP", wp, wp.vocab_trie.encoder), ("FastBPE", bpe, bpe._device_encoder())):
    tok.tokenize(lines[0])
    sc, pt, _, binding = enc._small_ctx
    datas = [l.encode() for l in lines]
    fn, h = sc._fn, sc.handle
    def timeit(f, reps=3):
        best = 1e9
        for _ in range(reps):
            t = time.perf_counter(); f(); best = min(best, time.perf_counter() - t)
        return 1e6 * best / len(lines)
    r = {}
    r["c_call_us"] = timeit(lambda: [fn(h, binding, d, len(d)) for d in datas])
    r["c_call_plus_copy_us"] = timeit(lambda: [sc.tokenize(binding, d) for d in datas])
    r["encode_text_us"] = timeit(lambda: [enc.encode_text(l) for l in lines])
    ids = [enc.encode_text(l) for l in lines]
    r["tokens_to_strs_us"] = timeit(lambda: [enc.tables.tokens_to_strs(i) for i in ids])
    r["tokenize_us"] = timeit(lambda: [tok.tokenize(l) for l in lines])
    r["empty_text_c_call_us"] = timeit(lambda: [fn(h, binding, b"", 0) for d in datas])
    cyc = []
    for d_ in datas[:200]:
        fn(h, binding, d_, len(d_)); cyc.append([int(sc._out[5]), int(sc._out[6]), int(sc._out[7])])
    import numpy as np
    r["kernel_cycles_pretok_encode_total_mean"] = np.mean(np.array(cyc), axis=0).tolist()
    one = datas[3]
    r["same_line_c_call_us"] = timeit(lambda: [fn(h, binding, one, len(one)) for d in datas])
    out[name] = r
print(json.dumps(out))
