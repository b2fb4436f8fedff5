"""Tokenize rate as the number of distinct word types grows (the word-type memo holds at most 2^22 entries): Zipf streams over
N synthetic types (bench_data.synth_type_table), FastWP with the pretrained 20 K vocabulary, ~400 MB per stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import bench, bench_data as BD
from subword_tokenizers_b200 import device, packing as P
from subword_tokenizers_b200.utils import naive_wp_encode_ids

dev = torch.device("cuda", 0)
tab = P.WpTables(bench.load_golden("pretrained_wp_vocab.json.gz"))
wenc = device.WpEncoder(tab, naive_wp_encode_ids("##", tab))
out = {}
for n_types in (20_000, 200_000, 2_000_000):
    mat, lens_ = BD.synth_type_table(n_types, 1)
    t_arena, t_off = BD.table_to_utf8(mat, lens_)
    rng = np.random.Generator(np.random.PCG64(2))
    n_words = 45_000_000
    w = 1.0 / np.arange(1, n_types + 1)                              # Zipf(s = 1) over the types
    cdf = np.cumsum(w); cdf /= cdf[-1]
    draw = np.searchsorted(cdf, rng.random(n_words)).astype(np.int64)
    d_tlen = torch.from_numpy(np.diff(t_off.astype(np.int64))).to(dev)
    d_toff = torch.from_numpy(t_off[:-1].astype(np.int64)).to(dev)
    d_tarena = torch.from_numpy(t_arena).to(dev)
    d_draw = torch.from_numpy(draw).to(dev)
    lens = d_tlen[d_draw]
    off = torch.zeros(n_words + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=off[1:])
    total = int(off[-1].item())
    src = torch.repeat_interleave(d_toff[d_draw] - off[:-1], lens) + torch.arange(total, device=dev)
    d_arena = d_tarena[src]
    d_off = off.to(torch.int32)
    del src, lens
    ws = torch.empty(device._lib.load().swt_encode_workspace_bytes(n_words, 0), dtype=torch.uint8, device=dev)
    cap = total + n_words + 16
    ids = torch.empty(cap, dtype=torch.int32, device=dev); tok = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
    status = torch.empty(8, dtype=torch.int32, device=dev)
    for _ in range(2):
        wenc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        wenc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    b.record(); torch.cuda.synchronize()
    nt, _ = wenc.check_status(status)
    st = status.cpu().numpy()
    ms = a.elapsed_time(b) / 3
    out[n_types] = {"MB": total / 1e6, "ms": ms, "GB_per_s": total / ms / 1e6, "tokens_per_word": nt / n_words, "memo_types": int(st[4]),
                    "distinct_types_in_stream": int(len(np.unique(draw)))}
    print(n_types, out[n_types], flush=True)
    del d_arena, d_off, ids, tok, ws, off, d_draw
print(json.dumps(out))
