"""Cost of one cross-GPU barrier of the trainer's peer exchange (P2P flag stores + polling), eager launches and inside a CUDA graph.
    torchrun --nproc-per-node N profiles/peer_barrier.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from subword_tokenizers_b200 import device, packing as P
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tt = P.TrainTypes(["abc", "abd", "bcd", "abcd"] * 8)
a, b = device.shard_types(tt.off, world)[rank]
eng = device.CudaTrainEngine(tt.syms[int(tt.off[a]):int(tt.off[b])], tt.off[a:b + 1] - tt.off[a], tt.freq[a:b], tt.n_alpha, 64, tt.n_alpha, 8, int(tt.off[a]), rank, world)
assert eng.setup_peer_exchange(), getattr(eng, "peer_error", "?")
out = {}
with torch.cuda.stream(eng.stream):
    for name, n in (("eager", 2000),):
        eng.lib.swt_bpe_train_exchange_probe(eng.handle, 50, eng._sp()); eng.stream.synchronize(); dist.barrier()
        t = time.perf_counter(); eng.lib.swt_bpe_train_exchange_probe(eng.handle, n, eng._sp()); eng.stream.synchronize(); out[name + "_us_per_barrier"] = 1e6 * (time.perf_counter() - t) / n
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=eng.stream):
        eng.lib.swt_bpe_train_exchange_probe(eng.handle, 500, eng._sp())
    eng.stream.synchronize(); dist.barrier()
    g.replay(); eng.stream.synchronize(); dist.barrier()
    t = time.perf_counter()
    for _ in range(4): g.replay()
    eng.stream.synchronize(); out["graph_us_per_barrier"] = 1e6 * (time.perf_counter() - t) / 2000
    # all-reduce of 16 bytes through NCCL for comparison
    x = torch.zeros(2, dtype=torch.int64, device=eng.dev)
    for _ in range(20): dist.all_reduce(x)
    eng.stream.synchronize(); t = time.perf_counter()
    for _ in range(500): dist.all_reduce(x)
    eng.stream.synchronize(); out["nccl_allreduce_16B_us"] = 1e6 * (time.perf_counter() - t) / 500
if rank == 0: print(json.dumps(out), flush=True)
eng.close(); dist.destroy_process_group()
