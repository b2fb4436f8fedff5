"""BASELINE config 3 driver: BPE training on synthetic word types (bench_data.synth_type_table), 1..8 GPUs.

    python profiles/train_scale.py --types 10000000 --max-vocab 32000 [--check-oracle-steps K]
    torchrun --nproc-per-node N profiles/train_scale.py ...        (per-step exchange over NVLink peer memory, NCCL for the initial counts)

Prints one JSON line (merges/s, sha256 of the merge list, tie steps, optional oracle-checked prefix)."""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_data as BD


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--types", type=int, default=1_000_000)
    ap.add_argument("--max-vocab", type=int, default=32000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--check-oracle-steps", type=int, default=0, help="compare the first K merges with the CPU oracle")
    ap.add_argument("--steps-per-sync", type=int, default=512)
    ap.add_argument("--no-peer", action="store_true", help="keep the two NCCL collectives per step (round-1 exchange)")
    ap.add_argument("--timing", action="store_true", help="per-kernel device times (eager launches, synchronised: slow)")
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from subword_tokenizers_b200 import device
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.no_peer:
        os.environ["SWT_NO_PEER_EXCHANGE"] = "1"
    if args.timing:
        device.tune("train_timing", 1)
    t0 = time.perf_counter()
    mat, lens = BD.synth_type_table(args.types, args.seed)
    cps, off = BD.table_to_cps(mat, lens)
    alpha = np.unique(cps)
    syms = np.searchsorted(alpha, cps).astype(np.uint32)
    freq = BD.zipf_freqs(args.types, args.seed)
    t_gen = time.perf_counter() - t0
    n_alpha = len(alpha)
    a, b = device.shard_types(off, world)[rank]
    max_len = int(lens.max())
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
           "merges": int(len(l)), "seconds": dt, "merges_per_s": len(l) / dt, "us_per_step": 1e6 * dt / max(1, len(l)), "vocab_size": int(state["vocab_size"]),
           "table_entries": int(state["n_table_entries"]), "table_cap": int(state["table_cap"]), "tie_steps": int(state["n_tie_steps"]), "tie_steps_listed": int(state["n_tie_listed"]),
           "live_slots_rank0": int(state["n_live_slots"]), "merges_sha256": h, "gen_seconds": t_gen, "halt": int(state["halt"]),
           "exchange": getattr(eng, "exchange_kind", "none"),
           "peer_barriers": int(state.get("n_peer_barriers", 0)), "peer_kernel_cycles_per_step": [int(x) / max(1, len(l)) for x in state.get("peer_kernel_cycles", [0, 0, 0])], "peer_wait_cycles_per_barrier": int(state.get("peer_wait_cycles", 0)) / max(1, int(state.get("n_peer_barriers", 0)))}
    if args.check_oracle_steps and rank == 0:
        import oracle
        k = args.check_oracle_steps
        ol, orr, on, oc, _ = oracle.bpe_train(syms, off, freq, n_alpha, n_alpha + k, max_merges=k + 16)
        m = min(k, len(l), len(ol))
        out["oracle_prefix_checked"] = m
        out["oracle_prefix_equal"] = bool(np.array_equal(l[:m], ol[:m]) and np.array_equal(r[:m], orr[:m]) and
                                          np.array_equal(n[:m], on[:m]) and np.array_equal(c[:m], oc[:m]))
    eng.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
