import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from subword_tokenizers_b200 import device, packing as P
from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
pre = make_hf_tokenizer().backend_tokenizer.pre_tokenizer
words = [w for s in bench.load_golden("train-5K.json.gz") for w, _ in pre.pre_tokenize_str(s.lower())]
tt = P.TrainTypes(words)
max_len = int(np.diff(tt.off.astype(np.int64)).max())
for _ in range(2):
    eng = device.CudaTrainEngine(tt.syms, tt.off, tt.freq, tt.n_alpha, 8000, tt.n_alpha, max_len, 0, 0, 1, record_cap=8192)
    torch.cuda.synchronize(); t = time.perf_counter()
    l, r, n, c, state = device.run_training_loop(eng, 1, steps_per_sync=1024)
    torch.cuda.synchronize(); print("bpe 5k->8000:", len(l), time.perf_counter() - t)
    eng.close()
