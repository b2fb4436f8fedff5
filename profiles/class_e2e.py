"""The class call on a large text: FastWP.tokenize / FastBPE.tokenize(text of ~100 MB) -> List[str], wall time and where it goes
(device part vs the id -> string materialisation on the host)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from subword_tokenizers_b200 import FastWP, FastBPE, make_hf_tokenizer
from subword_tokenizers_b200.utils import WPTrie_E2E

arena, off = bench.ZipfStream.train5k(0).host_sample(12_000_000)
text = " ".join(arena[int(off[k]):int(off[k + 1])].tobytes().decode() for k in range(len(off) - 1))
hf = make_hf_tokenizer()
wp = FastWP(hf); wp.vocab = set(bench.load_golden("ref_wp_train5k_v8000_vocab.json.gz")); wp.vocab_trie = WPTrie_E2E(wp.vocab)
bpe = FastBPE(hf); bpe.merges_list = [tuple(p) for p in bench.load_golden("ref_bpe_train5k_v8000_merges.json.gz")]; bpe._rebuild_ranks()
out = {"text_bytes": len(text.encode()), "words": len(off) - 1}
for name, tok, enc in (("FastWP", wp, wp.vocab_trie.encoder), ("FastBPE", bpe, bpe._device_encoder())):
    tok.tokenize(text[:1_000_000])
    torch.cuda.synchronize(); t = time.perf_counter(); toks = tok.tokenize(text); t_all = time.perf_counter() - t
    t = time.perf_counter(); ids = enc.encode_text(text); t_ids = time.perf_counter() - t
    t = time.perf_counter(); strs = enc.tables.tokens_to_strs(ids); t_str = time.perf_counter() - t
    assert strs == toks
    out[name] = {"tokenize_s": t_all, "MB_per_s": out["text_bytes"] / t_all / 1e6, "encode_text_s (utf-8 encode, H2D, kernels, D2H of ids)": t_ids,
                 "tokens_to_strs_s": t_str, "tokens": len(toks)}
print(json.dumps(out))
