"""BASELINE config 3 driver: BPE training on synthetic word types, 1..8 GPUs.

    python profiles/train_scale.py --types 10000000 --max-vocab 32000 [--check-oracle-steps K]
    torchrun --nproc-per-node N profiles/train_scale.py ...        (NCCL pair-count all-reduce per merge step)

Word types: concatenations of 1-4 units drawn from the merged strings of the reference's pretrained BPE model plus
single letters of the train-5K alphabet, duplicates rejected, clipped to 22 characters; frequency of rank r is
max(1, floor(C / r)); the type order (which drives the tie-break) is a seeded permutation.  Prints one JSON line.
"""
import argparse, gzip, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_golden(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name), "rt", encoding="utf-8") as f:
        return json.load(f)


def synth_types(n_types, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    merges = load_golden("pretrained_bpe_merges.json.gz")
    units = sorted({a + b for a, b in merges})
    alphabet = sorted({c for u in units for c in u})
    units = units + alphabet
    nu = len(units)
    types, seen = [], set()
    while len(types) < n_types:
        m = max(1 << 16, (n_types - len(types)) * 5 // 4)
        k = rng.integers(1, 5, size=m)
        idx = rng.integers(0, nu, size=(m, 4))
        tgt = np.clip(np.rint(rng.normal(8.2, 3.0, size=m)), 2, 22).astype(np.int64)   # train-5K: mode 7-8, mean 8.18, max 22
        for i in range(m):
            w = "".join(units[j] for j in idx[i, :k[i]])[:tgt[i]]
            if w not in seen:
                seen.add(w); types.append(w)
                if len(types) == n_types:
                    break
    order = rng.permutation(n_types)                       # rank r -> position order[r]
    freq = np.zeros(n_types, dtype=np.int64)
    C = 20 * n_types
    freq[order] = np.maximum(1, C // np.arange(1, n_types + 1))
    return types, freq


def pack(types):
    from subword_tokenizers_b200 import packing as P
    cps, off = P.pack_strings_as_cps(types)
    alpha = np.unique(cps)
    syms = np.searchsorted(alpha, cps).astype(np.uint32)
    return syms, off, [chr(c) for c in alpha]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--types", type=int, default=1_000_000)
    ap.add_argument("--max-vocab", type=int, default=32000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--check-oracle-steps", type=int, default=0, help="compare the first K merges with the CPU oracle")
    ap.add_argument("--steps-per-sync", type=int, default=512)
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from subword_tokenizers_b200 import device
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    types, freq = synth_types(args.types, args.seed)
    syms, off, alphabet = pack(types)
    t_gen = time.perf_counter() - t0
    n_alpha = len(alphabet)
    a, b = device.shard_types(off, world)[rank]
    max_len = int(np.diff(off.astype(np.int64)).max())
    eng = device.CudaTrainEngine(syms[int(off[a]):int(off[b])], off[a:b + 1] - off[a], freq[a:b], n_alpha, args.max_vocab, n_alpha,
                                 max_len, int(off[a]), rank, world, record_cap=8192)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tic = time.perf_counter()
    l, r, n, c, state = device.run_training_loop(eng, world, steps_per_sync=args.steps_per_sync)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - tic
    h = hashlib.sha256(np.stack([l, r, n]).tobytes() + c.tobytes()).hexdigest()
    out = {"n_gpus": world, "n_types": args.types, "n_symbols": int(off[-1]), "n_alpha": n_alpha, "max_vocab": args.max_vocab,
           "merges": int(len(l)), "seconds": dt, "merges_per_s": len(l) / dt, "vocab_size": int(state["vocab_size"]),
           "table_entries": int(state["n_table_entries"]), "table_cap": int(state["table_cap"]),
           "live_slots_rank0": int(state["n_live_slots"]), "merges_sha256": h, "gen_seconds": t_gen, "halt": int(state["halt"])}
    if args.check_oracle_steps and rank == 0:
        import oracle
        k = args.check_oracle_steps
        ol, orr, on, oc, _ = oracle.bpe_train(syms, off, freq, n_alpha, n_alpha + k)
        m = min(k, len(l), len(ol))
        out["oracle_prefix_checked"] = m
        out["oracle_prefix_equal"] = bool(np.array_equal(l[:m], ol[:m]) and np.array_equal(r[:m], orr[:m]) and
                                          np.array_equal(n[:m], on[:m]) and np.array_equal(c[:m], oc[:m]))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
