"""Profiling driver: runs the FastWP and FastBPE encode kernels a few times over a Zipf stream (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from subword_tokenizers_b200 import device, packing as P
from subword_tokenizers_b200.utils import naive_wp_encode_ids

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
stream = bench.ZipfStream.train5k(0)
d_arena, d_off, n_words, off32 = stream.device_stream(nbytes, dev)
tab = P.WpTables(bench.load_golden("ref_wp_train5k_v8000_vocab.json.gz"))
wenc = device.WpEncoder(tab, naive_wp_encode_ids("##", tab))
benc = device.BpeEncoder(P.BpeTables([tuple(p) for p in bench.load_golden("ref_bpe_train5k_v8000_merges.json.gz")]))
ws = torch.empty(device._lib.load().swt_encode_workspace_bytes(n_words, 0), dtype=torch.uint8, device=dev)
cap = int(d_arena.numel()) + n_words + 16
ids = torch.empty(cap, dtype=torch.int32, device=dev)
tok = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
status = torch.empty(8, dtype=torch.int32, device=dev)
only = sys.argv[3] if len(sys.argv) > 3 else ""
for name, enc in (("wp", wenc), ("bpe", benc)):
    if only and name != only:
        continue
    for _ in range(reps):
        enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status); b.record(); torch.cuda.synchronize()
    nt, h6 = enc.check_status(status)
    st = status.cpu().numpy()
    print(name, "ms", a.elapsed_time(b), "tokens", nt, "memo types", int(st[4]), "slow-path words", int(st[5]), "words", n_words, "bytes", int(d_arena.numel()))
