"""Round-2 kernel experiments on one GPU: the encode call on the bench stream and on many-type streams under the swt_tune knobs,
every variant checked against the CPU oracle on a prefix.  Prints one JSON line per measurement.

    python profiles/exp_r2.py [--bytes 1000000000] [--many 2000000] [--skip-many]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_data as B
from subword_tokenizers_b200 import device, packing as P
from subword_tokenizers_b200.utils import naive_wp_encode_ids

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=1_000_000_000)
ap.add_argument("--many", type=int, default=2_000_000)
ap.add_argument("--many-words", type=int, default=45_000_000)
ap.add_argument("--skip-many", action="store_true")
ap.add_argument("--quick", action="store_true", help="default knobs only")
ap.add_argument("--check-words", type=int, default=2_000_000)
args = ap.parse_args()
dev = torch.device("cuda", 0)
lib = device._lib.load()


def oracle_ids(kind, tab, arena, off):
    import oracle
    if kind == "wp":
        alnum, space = P.unicode_class_bitmaps()
        ids, _, _ = oracle.WpTrie(tab, alnum).encode(arena, off, space)
        return ids
    ids, _ = oracle.bpe_encode(tab, arena, off)
    return ids


def run(name, enc, kind, tab, d_arena, d_off, n_words, host_prefix, knobs, reps=5):
    for k, v in knobs.items():
        device.tune(k, v)
    n_bytes = int(d_arena.numel())
    ws = torch.empty(lib.swt_encode_workspace_bytes(n_words, 0), dtype=torch.uint8, device=dev)
    cap = n_bytes + n_words + 16
    ids = torch.empty(cap, dtype=torch.int32, device=dev); tok = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
    status = torch.empty(8, dtype=torch.int32, device=dev)
    for _ in range(2):
        enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    device.tune("timing", 1)
    enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    torch.cuda.synchronize()
    device.tune("timing", 0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    b.record(); torch.cuda.synchronize()
    nt, h6 = enc.check_status(status)
    st = status.cpu().numpy()
    ms = a.elapsed_time(b) / reps
    alg = n_bytes + 8 * (n_words + 1) + 4 * nt
    # parity on the prefix
    h_arena, h_off = host_prefix
    npre = len(h_off) - 1
    want = oracle_ids(kind, tab, h_arena, h_off)
    tk = tok[:npre + 1].cpu().numpy().view(np.uint32)
    got = ids[:int(tk[-1])].cpu().numpy().view(np.uint32)
    ok = bool(len(got) == len(want) and np.array_equal(got, want))
    out = {"name": name, "knobs": knobs, "ms": round(ms, 4), "GB_per_s": round(n_bytes / ms / 1e6, 1), "roofline_frac": round(alg / ms / 1e6 / 6545.3, 4),
           "tokens_per_word": round(nt / n_words, 3), "memo_types": int(st[4]), "slow_words": int(st[5]), "parity_prefix_words": npre, "parity": ok}
    print(json.dumps(out), flush=True)
    for k in knobs:                       # back to defaults
        device.tune(k, {"memo_max_log2": 22, "memo_off": 0, "bulk_store": 1, "warp_words": 3, "bpe_queue": 1}[k])
    del ws, ids, tok
    return out


# ---- bench stream (train-5K types, 22,971)
stream = B.ZipfStream.train5k(0)
d_arena, d_off, n_words, off32 = stream.device_stream(args.bytes, dev)
pre = B.ZipfStream.train5k(0).host_sample(min(args.check_words, n_words))
vocab = B.load_golden("ref_wp_train5k_v8000_vocab.json.gz")
wtab = P.WpTables(vocab)
wenc = device.WpEncoder(wtab, naive_wp_encode_ids("##", wtab))
merges = [tuple(p) for p in B.load_golden("ref_bpe_train5k_v8000_merges.json.gz")]
btab = P.BpeTables(merges)
benc = device.BpeEncoder(btab)
for knobs in (({},) if args.quick else ({}, {"bulk_store": 0}, {"memo_max_log2": 20}, {"memo_max_log2": 18}, {"memo_off": 1})):
    run("wp_bench_stream", wenc, "wp", wtab, d_arena, d_off, n_words, pre, knobs, reps=5 if not knobs.get("memo_off") else 2)
for knobs in (({}, {"bpe_queue": 0}) if args.quick else ({}, {"bulk_store": 0}, {"bpe_queue": 0}, {"bpe_queue": 0, "warp_words": 0}, {"warp_words": 0}, {"warp_words": 8}, {"memo_off": 1})):
    run("bpe_bench_stream", benc, "bpe", btab, d_arena, d_off, n_words, pre, knobs, reps=5 if not knobs.get("memo_off") else 2)
del d_arena, d_off

# ---- many-type stream: Zipf over N synthetic types (the tail beyond rank C has weight 1), pretrained 20K models
if not args.skip_many:
    mat, lens = B.synth_type_table(args.many, 1)
    t_arena, t_off = B.table_to_utf8(mat, lens)
    ms_ = B.ZipfStream(t_arena, t_off, 200_000, 2)
    target = int(args.many_words * ms_.mean_len)
    d_arena, d_off, n_words, off32 = ms_.device_stream(target, dev)
    pre = B.ZipfStream(t_arena, t_off, 200_000, 2).host_sample(min(args.check_words, n_words))
    wtab2 = P.WpTables(B.load_golden("pretrained_wp_vocab.json.gz"))
    wenc2 = device.WpEncoder(wtab2, naive_wp_encode_ids("##", wtab2))
    btab2 = P.BpeTables([tuple(p) for p in B.load_golden("pretrained_bpe_merges.json.gz")])
    benc2 = device.BpeEncoder(btab2)
    for knobs in (({},) if args.quick else ({}, {"memo_max_log2": 23}, {"memo_max_log2": 20}, {"memo_off": 1})):
        run("wp_many_types", wenc2, "wp", wtab2, d_arena, d_off, n_words, pre, knobs, reps=3)
    for knobs in (({},) if args.quick else ({}, {"memo_max_log2": 23}, {"bpe_queue": 0}, {"memo_off": 1})):
        run("bpe_many_types", benc2, "bpe", btab2, d_arena, d_off, n_words, pre, knobs, reps=3)
