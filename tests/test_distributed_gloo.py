"""world_size-2 gloo test of the N>1 training path on the CPU: the product's loop (run_training_loop:
all_reduce of the initial counts, all_gather of tie-break candidates, all_reduce of pair-count deltas) and
its sharding (shard_types) driven with a numpy rank engine; both ranks must reproduce the oracle's merges."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _corpora():
    rng = np.random.default_rng(42)
    out = []
    for alpha, n_words, max_vocab in (("ab", 40, 12), ("abc", 200, 40), ("abcdefgh", 600, 120), ("aab", 50, 30)):
        words = ["".join(rng.choice(list(alpha), size=int(rng.integers(1, 9)))) for _ in range(n_words)]
        words += ["a" * 9, "ab" * 5]
        out.append((words, max_vocab))
    kat = load_golden("kat_tests_resources.json")
    out.append(([w.strip(".").lower() for s in kat["corpus"] for w in s.split()], 25))
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from numpy_train_engine import NumpyTrainEngine
    from subword_tokenizers_b200 import packing as P
    from subword_tokenizers_b200.device import run_training_loop, shard_types
    results = []
    for words, max_vocab in _corpora():
        tt = P.TrainTypes(words)
        t0, t1 = shard_types(tt.off, world)[rank]
        off = tt.off[t0:t1 + 1] - tt.off[t0]
        eng = NumpyTrainEngine(tt.syms[int(tt.off[t0]):int(tt.off[t1])], off, tt.freq[t0:t1], tt.n_alpha, max_vocab,
                               tt.n_alpha, int(tt.off[t0]), rank, world, record_cap=5)
        l, r, n, c, state = run_training_loop(eng, world, steps_per_sync=3)
        results.append((l.tolist(), r.tolist(), n.tolist(), c.tolist(), state["vocab_size"]))
    q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_training_matches_oracle():
    import oracle
    from subword_tokenizers_b200 import packing as P
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, (words, max_vocab) in enumerate(_corpora()):
        tt = P.TrainTypes(words)
        l, r, n, c, vs = oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab)
        expect = (l.tolist(), r.tolist(), n.tolist(), c.tolist(), vs)
        assert got[0][k] == expect, "rank 0, corpus %d" % k
        assert got[1][k] == expect, "rank 1, corpus %d" % k


def test_single_rank_numpy_engine_matches_oracle():
    """The same engine on one rank: pins the protocol arithmetic (delta vectors L/R/ZZ/M) to the oracle."""
    import oracle
    from numpy_train_engine import NumpyTrainEngine
    from subword_tokenizers_b200 import packing as P
    from subword_tokenizers_b200.device import run_training_loop
    for words, max_vocab in _corpora():
        tt = P.TrainTypes(words)
        eng = NumpyTrainEngine(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab, tt.n_alpha, 0, 0, 1, record_cap=7)
        l, r, n, c, state = run_training_loop(eng, 1, steps_per_sync=4)
        ol, orr, on, oc, vs = oracle.bpe_train(tt.syms, tt.off, tt.freq, tt.n_alpha, max_vocab)
        assert (l.tolist(), r.tolist(), n.tolist(), c.tolist()) == (ol.tolist(), orr.tolist(), on.tolist(), oc.tolist())
        assert state["vocab_size"] == vs
