"""Worker of test_multi_gpu_trainer_matches_single_gpu_and_oracle: launched with torch.distributed.run, one process per GPU.
Every rank trains its shard of the word types with CudaTrainEngine (per-step exchange over NCCL / peer memory); rank 0 also trains
the whole table on one GPU and runs the CPU oracle; all merge lists must be identical."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench_data as BD
    from subword_tokenizers_b200 import device, packing as P
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(11)
    alphabet = list("abcdefghijklmnopqrstuvwxyz")
    cases = []
    # 1. tie-heavy: every type has frequency 1, tiny table (several rehashes)
    words = ["".join(rng.choice(alphabet, size=int(rng.integers(2, 12)))) for _ in range(4000)]
    tt = P.TrainTypes(words)
    cases.append(("ties", tt.syms, tt.off, tt.freq, tt.n_alpha, 600, 1024, 600))
    # 2. 200 k synthetic types with Zipf frequencies
    mat, lens = BD.synth_type_table(200_000, 5)
    cps, off = BD.table_to_cps(mat, lens)
    alpha = np.unique(cps)
    syms = np.searchsorted(alpha, cps).astype(np.uint32)
    cases.append(("zipf200k", syms, off, BD.zipf_freqs(200_000, 5), len(alpha), len(alpha) + 3000, 0, 200))
    # 3. alphabet beyond the dense-count limit (sparse exchange of the initial counts)
    cjk = np.arange(0x4E00, 0x4E00 + 4500)
    words = ["".join(chr(c) for c in rng.choice(cjk, size=int(rng.integers(1, 6)))) for _ in range(20_000)]
    tt3 = P.TrainTypes(words)
    cases.append(("cjk", tt3.syms, tt3.off, tt3.freq, tt3.n_alpha, tt3.n_alpha + 200, 0, 200))
    for name, syms, off, freq, n_alpha, max_vocab, table_cap, oracle_steps in cases:
        off = np.asarray(off, dtype=np.uint64)
        a, b = device.shard_types(off, world)[rank]
        max_len = int(np.diff(off.astype(np.int64)).max())
        cap = table_cap or (2 * int(off[-1]) + 8 * max_vocab if n_alpha > 4096 else 0)
        eng = device.CudaTrainEngine(syms[int(off[a]):int(off[b])], off[a:b + 1] - off[a], freq[a:b], n_alpha, max_vocab, n_alpha, max_len,
                                     int(off[a]), rank, world, record_cap=256, table_cap=cap)
        l, r, n, c, state = device.run_training_loop(eng, world, steps_per_sync=256)
        eng.close()
        sha = hashlib.sha256(np.stack([l, r, n]).tobytes() + c.tobytes()).hexdigest()
        h = torch.tensor([int(sha[:15], 16)], dtype=torch.int64, device=dev)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        assert all(int(x.item()) == int(h.item()) for x in hs), "%s: ranks disagree" % name
        if rank == 0:
            eng1 = device.CudaTrainEngine(syms, off, freq, n_alpha, max_vocab, n_alpha, max_len, 0, 0, 1, record_cap=256, table_cap=table_cap)
            l1, r1, n1, c1, s1 = device.run_training_loop(eng1, 1, steps_per_sync=256)
            eng1.close()
            assert len(l) == len(l1) > 0, "%s: %d vs %d merges" % (name, len(l), len(l1))
            assert np.array_equal(l, l1) and np.array_equal(r, r1) and np.array_equal(n, n1) and np.array_equal(c, c1), "%s: N-GPU != 1-GPU" % name
            import oracle
            ol, orr, on, oc, _ = oracle.bpe_train(syms, off, freq, n_alpha, min(max_vocab, n_alpha + oracle_steps))   # a prefix of the merges
            m = min(len(ol), len(l))
            assert m > 0 and np.array_equal(l[:m], ol[:m]) and np.array_equal(r[:m], orr[:m]) and np.array_equal(n[:m], on[:m]) and \
                np.array_equal(c[:m], oc[:m]), "%s: != oracle" % name
            print("case %s: %d merges, %d checked against the oracle, exchange=%s" % (name, len(l), m, getattr(eng, "exchange_kind", "nccl")), flush=True)
        dist.barrier()
    if rank == 0:
        print("MGPU_TRAIN_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
