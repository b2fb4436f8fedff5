"""Wall time of the whole train(corpus) call: pre-tokenization + word-type counting + merge loop.
Host pre-processing (Rust BertPreTokenizer + Counter, what the reference does) vs the device pre-processing."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from subword_tokenizers_b200 import NaiveBPE, NaiveWP, make_hf_tokenizer, device, packing as P

corpus = bench.load_golden("train-5K.json.gz")
hf = make_hf_tokenizer()
out = {}
for reps in (1, 40):
    c = corpus * reps
    nbytes = sum(len(s.encode()) for s in c)
    nb = NaiveBPE(hf); nb.train(c[:100], 100)                                   # warm-up (tables, library)
    torch.cuda.synchronize(); t = time.perf_counter()
    words = nb._pre_tokenized_words(c); types_h = P.TrainTypes(words)
    t_host = time.perf_counter() - t
    t = time.perf_counter(); types_d = device.device_train_types(c, wordpiece=False); torch.cuda.synchronize()
    t_dev = time.perf_counter() - t
    assert types_d.alphabet == types_h.alphabet and (types_d.freq == types_h.freq).all() and (types_d.syms == types_h.syms).all()
    t = time.perf_counter(); nb.train(c, 8000); torch.cuda.synchronize(); t_train = time.perf_counter() - t
    nw = NaiveWP(hf)
    t = time.perf_counter(); nw.train(c, 8000); torch.cuda.synchronize(); t_wp = time.perf_counter() - t
    out["x%d" % reps] = {"corpus_MB": nbytes / 1e6, "words": len(words), "types": types_h.n_types,
                         "host_preprocessing_s": t_host, "device_preprocessing_s": t_dev,
                         "NaiveBPE.train_total_s": t_train, "merge_loop_s": nb.last_train_stats["merge_loop_s"], "merges": len(nb.merges_list),
                         "NaiveWP.train_total_s": t_wp}
print(json.dumps(out))
