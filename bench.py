#!/usr/bin/env python
"""bench.py -- headline benchmark of the tokenize/train hot paths (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bytes B] [--no-also]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): FastWP, vocabulary of 8000 trained by the UNMODIFIED reference on
data/train-5K.json (data/train-85k.json is missing from the reference checkout; fixture
tests/golden/ref_wp_train5k_v8000_vocab.json.gz), tokenizing a 1 GB synthetic Zipf word stream per GPU.
A "step" is one pass of the FastWP encode call over the whole stream.

  value         MB/s (10^6 input arena bytes per second) with the stream resident in HBM, CUDA-event timed, max over ranks, whole job
  e2e           same metric through the host-buffer C ABI from RAW TEXT (swt_tokenize_text_host): pinned host text in, lower-casing
                + whitespace split + encode on the device, flat 16-bit token ids out; H2D and D2H inside the timed region; with the
                measured ceiling of the same copies without any kernel (frac_of_copy_ceiling)
  roofline      algorithmic bytes (arena + 4 B/word offset read, 4 B/token + 4 B/word offset written) per call / mean call time,
                against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  parity        the ids of a prefix of THIS run's stream are compared with the CPU oracle inside the bench (parity_checked_words)
  cpu_baseline  the C oracle port of the reference's FastWP path on the host cores, bounded sample (N = 1)
  also          at EVERY N: FastBPE tokenize over the same stream; BPE training on 10 M synthetic word types sharded over the N
                GPUs (NCCL / peer exchange per merge step; configs[2]); the many-type 32 K-vocabulary workload (configs[3] shape).
                At N = 1 additionally: device pre-tokenizer, train-5K trainings.

--impl reference: times the CPU implementation of the same path (oracle C port -- the reference itself is pure Python and cannot
travel to the GPU box) with all host threads on a bounded sample (a prefix) of the same stream.
"""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench_data as BD                                    # noqa: E402
from bench_data import load_golden, ZipfStream             # noqa: E402,F401

ZIPF_C = 200_000
KERNELS_PER_ENCODE_CALL = 7     # memo_clear, count, long_count, scan_groups, scan_top, emit, long_emit


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML from a background thread (every ~2 ms),
    so that even a timed region of a few tens of milliseconds gets samples DURING the region."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        import threading
        self.ok = False
        self.sm, self.reasons, self.max_sm = [], set(), None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:                       # noqa: BLE001
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:                        # noqa: BLE001
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self._stop.set()
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(key):
    """DRAM bytes per call from this round's ncu --set full capture (profiles/r02_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(path)).get(key)
    except Exception:
        return None


def roofline(alg_bytes, kern_ms, kernel, traffic_key=None):
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    t = measured_traffic(traffic_key) if traffic_key else None
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (t or {}).get("bytes") if t else None, "traffic_source": (t or {}).get("source") if t else None,
            "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": kern_ms}


def device_text(d_arena, d_off, n_words):
    """Raw text on the device for the pre-tokenizer: the stream's words joined (and followed) by single spaces.
    -> (uint8 tensor padded to a multiple of 4, n_text_bytes)."""
    import torch
    dev = d_arena.device
    n_bytes = int(d_arena.numel())
    off = d_off[: n_words + 1].long()
    wid = torch.repeat_interleave(torch.arange(n_words, device=dev, dtype=torch.int32), off[1:] - off[:-1])
    pos = torch.arange(n_bytes, device=dev, dtype=torch.int64)
    pos += wid
    del wid
    n_text = n_bytes + n_words
    text = torch.full(((n_text + 7) // 4 * 4,), 0x20, dtype=torch.uint8, device=dev)
    text[pos] = d_arena
    text[n_text:] = 0
    return text, n_text


# ------------------------------------------------------------------------------------------------------------
# CPU oracle (the checker, and the cpu_baseline / reference arm)
# ------------------------------------------------------------------------------------------------------------
def oracle_encode(kind, tab, arena, off, threads):
    """Oracle ids of the words (arena, off) on `threads` host threads (ctypes releases the GIL; disjoint word slices).
    -> (ids u32 concatenated in word order, seconds)."""
    import oracle
    from subword_tokenizers_b200 import packing as P
    n = len(off) - 1
    threads = max(1, min(threads, n // 1000 + 1))
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    slices = []
    for t in range(threads):
        a, b = int(bounds[t]), int(bounds[t + 1])
        slices.append((arena[int(off[a]):int(off[b])].copy(), (off[a:b + 1] - off[a]).astype(np.uint64)))
    if kind == "wp":
        alnum, space = P.unicode_class_bitmaps()
        tries = [oracle.WpTrie(tab, alnum) for _ in range(threads)]
        work = lambda t: tries[t].encode(slices[t][0], slices[t][1], space)[0]       # noqa: E731
    else:
        work = lambda t: oracle.bpe_encode(tab, slices[t][0], slices[t][1])[0]       # noqa: E731
    tic = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, range(threads)))
    sec = time.perf_counter() - tic
    return np.concatenate(parts) if parts else np.zeros(0, np.uint32), sec


def python_reference_rate(vocab, arena, off, max_bytes=3_000_000):
    """The UNMODIFIED Python reference (baseline/_ref/source, vendored by __graft_entry__.build()) on ONE host core: FastWP.tokenize over
    the text form of a prefix of the stream.  -> dict or None when the reference is not vendored."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_root, "source", "wordpiece.py")):
        return None
    try:
        sys.path.insert(0, ref_root)
        from source.wordpiece import FastWP as RefFastWP            # type: ignore
        from source.utils import WPTrie_E2E as RefTrie             # type: ignore
        from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
        n = int(np.searchsorted(off, max_bytes))
        words = [arena[int(off[k]):int(off[k + 1])].tobytes().decode("utf-8", "surrogatepass") for k in range(n)]
        text = " ".join(words)
        tok = RefFastWP(make_hf_tokenizer())
        tok.vocab = set(vocab)
        tok.vocab_trie = RefTrie(tok.vocab)
        tok.tokenize(text[:2000])
        tic = time.perf_counter()
        out = tok.tokenize(text)
        sec = time.perf_counter() - tic
        nbytes = int(off[n])
        return {"value": nbytes / sec / 1e6, "unit": "MB/s", "cores": 1, "kind": "reference",
                "sample": "reference FastWP.tokenize (pure Python, source/wordpiece.py:233-316) on the first %d words (%.1f MB) of the stream, "
                          "%.1f s, %d tokens" % (n, nbytes / 1e6, sec, len(out))}
    except Exception as e:                                          # noqa: BLE001
        return {"error": repr(e)}
    finally:
        if ref_root in sys.path:
            sys.path.remove(ref_root)


def check_prefix(kind, tab, host_prefix, d_ids, d_tok, threads):
    """Compares the GPU ids of the stream's first words with the oracle. -> (n_words_checked, equal, oracle_seconds, bytes)."""
    h_arena, h_off = host_prefix
    npre = len(h_off) - 1
    want, sec = oracle_encode(kind, tab, h_arena, h_off, threads)
    tk = d_tok[:npre + 1].cpu().numpy().view(np.uint32)
    got = d_ids[:int(tk[-1])].cpu().numpy().view(np.uint32)
    return npre, bool(len(got) == len(want) and np.array_equal(got, want)), sec, int(len(h_arena))


# ------------------------------------------------------------------------------------------------------------
def time_encode(enc, d_arena, d_off, n_words, steps, warmup, dist, world):
    """Device-resident encode calls, CUDA-event timed. -> dict(ms per step list, total ms (max over ranks), tokens, h6, buffers)."""
    import torch
    from subword_tokenizers_b200 import device
    dev = d_arena.device
    n_bytes = int(d_arena.numel())
    lib_ws = device._lib.load().swt_encode_workspace_bytes(n_words, 0)
    d_ws = torch.empty(lib_ws, dtype=torch.uint8, device=dev)
    out_cap = n_bytes + n_words + 16
    d_ids = torch.empty(out_cap, dtype=torch.int32, device=dev)
    d_tok = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
    d_status = torch.empty(8, dtype=torch.int32, device=dev)

    def one_pass():
        enc.encode_into(d_arena, d_off, n_words, 0, d_ids, out_cap, d_tok, d_ws, d_status)
    for _ in range(max(3, warmup)):
        one_pass()
    n_tokens, h6 = enc.check_status(d_status)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(steps):
        one_pass()
        ev[k + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    total_ms = ev[0].elapsed_time(ev[steps])
    st = d_status.cpu().numpy().astype(np.uint32)
    enc.check_status(d_status)
    return {"step_ms": step_ms, "total_ms": total_ms, "n_tokens": n_tokens, "h6": h6, "d_ids": d_ids, "d_tok": d_tok,
            "memo_types": int(st[4]), "slow_words": int(st[5]), "ws": d_ws, "status": d_status, "out_cap": out_cap}


def reduce_max_sum(dist, world, dev, ms, nbytes):
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    b = torch.tensor([float(nbytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(t.item()), float(b.item())


def copy_ceiling(dev, h2d_bytes, d2h_bytes, batch_bytes, steps, dist, world):
    """The same host<->device copies as one e2e step (batches of `batch_bytes` of input, five slots / streams, H2D of batch k+1 and
    D2H of batch k-1 in flight together) without any kernel: the PCIe ceiling of the e2e number. -> seconds per step (max over ranks)."""
    import torch
    n_batches = max(1, (h2d_bytes + batch_bytes - 1) // batch_bytes)
    out_per = (d2h_bytes + n_batches - 1) // n_batches
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(out_per * n_batches, dtype=torch.uint8).pin_memory()
    slots = [(torch.empty(batch_bytes, dtype=torch.uint8, device=dev), torch.empty(out_per, dtype=torch.uint8, device=dev),
              torch.cuda.Stream(device=dev)) for _ in range(5)]

    def one():
        for k in range(n_batches):
            d_in, d_out, st = slots[k % 5]
            a, b = k * batch_bytes, min(h2d_bytes, (k + 1) * batch_bytes)
            with torch.cuda.stream(st):
                d_in[: b - a].copy_(h_in[a:b], non_blocking=True)
                h_out[k * out_per:(k + 1) * out_per].copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    one()
    if world > 1:
        dist.barrier()
    tic = time.perf_counter()
    for _ in range(steps):
        one()
    sec = (time.perf_counter() - tic) / steps
    t = torch.tensor([sec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------
def cached_synth_table(n_types, seed, rank, dist, world):
    """synth_type_table, generated once per box (rank 0) and shared through the temp directory."""
    path = os.path.join(tempfile.gettempdir(), "swt_synth_%d_%d.npz" % (n_types, seed))
    if rank == 0 and not os.path.isfile(path):
        mat, lens = BD.synth_type_table(n_types, seed)
        tmp = path + ".%d.tmp.npz" % os.getpid()
        np.savez(tmp, mat=mat, lens=lens)
        os.replace(tmp, path)
    if world > 1:
        dist.barrier()
    z = np.load(path)
    return z["mat"], z["lens"]


def train_10m(args, dev, rank, world, dist):
    """BASELINE configs[2]: BPE training on N_TYPES synthetic word types, max_vocab 32000, types sharded over the ranks."""
    import torch
    from subword_tokenizers_b200 import device
    tic = time.perf_counter()
    mat, lens = cached_synth_table(args.train_types, 0, rank, dist, world)
    cps, off = BD.table_to_cps(mat, lens)
    alpha = np.unique(cps)
    syms = np.searchsorted(alpha, cps).astype(np.uint32)
    freq = BD.zipf_freqs(args.train_types, 0)
    n_alpha = len(alpha)
    t_gen = time.perf_counter() - tic
    a, b = device.shard_types(off, world)[rank]
    max_len = int(lens.max())
    eng = device.CudaTrainEngine(syms[int(off[a]):int(off[b])], off[a:b + 1] - off[a], freq[a:b], n_alpha, args.train_vocab, n_alpha,
                                 max_len, int(off[a]), rank, world, record_cap=8192)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tic = time.perf_counter()
    l, r, n, c, state = device.run_training_loop(eng, world, steps_per_sync=512)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - tic
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    sha = hashlib.sha256(np.stack([l, r, n]).tobytes() + c.tobytes()).hexdigest()
    same = True
    if world > 1:                                          # every rank must hold the same merge list
        h = torch.tensor([int(sha[:15], 16)], dtype=torch.int64, device=dev)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        same = all(int(x.item()) == int(h.item()) for x in hs)
    out = {"value": len(l) / dt, "unit": "merges/s", "n_gpus": world, "merges": int(len(l)), "seconds": dt, "n_types": int(args.train_types),
           "n_symbols": int(off[-1]), "n_alpha": n_alpha, "max_vocab": args.train_vocab, "merges_sha256": sha,
           "all_ranks_same_merges": same, "table_entries": int(state["n_table_entries"]), "table_cap": int(state["table_cap"]),
           "exchange": getattr(eng, "exchange_kind", "nccl all_gather + all_reduce per step"), "data_gen_seconds": t_gen,
           "workload": "BASELINE configs[2]: %d synthetic word types (SURVEY.md 8d: units of the pretrained merges, length histogram of "
                       "train-5K, Zipf frequencies), max_vocab %d, types sharded over %d GPU(s); merge loop only" % (args.train_types, args.train_vocab, world)}
    if rank == 0 and args.check_oracle_steps:
        import oracle
        k = args.check_oracle_steps
        tic = time.perf_counter()
        ol, orr, on, oc, _ = oracle.bpe_train(syms, off, freq, n_alpha, n_alpha + k)
        m = min(k, len(l), len(ol))
        out["oracle_prefix_checked"] = int(m)
        out["oracle_prefix_equal"] = bool(m > 0 and np.array_equal(l[:m], ol[:m]) and np.array_equal(r[:m], orr[:m]) and
                                          np.array_equal(n[:m], on[:m]) and np.array_equal(c[:m], oc[:m]))
        out["cpu_reference_merges_per_s"] = m / (time.perf_counter() - tic)          # the oracle = the reference's full recount per merge
    eng.close()
    strs = [chr(int(x)) for x in alpha]
    merges = []
    for x, y, z in zip(l.tolist(), r.tolist(), n.tolist()):
        merges.append((strs[x], strs[y]))
        if z == len(strs):
            strs.append(strs[x] + strs[y])
    return out, merges, strs, n_alpha


def many_types(args, dev, rank, world, dist, merges, strs, n_alpha):
    """BASELINE configs[3] shape: FastBPE + FastWP over a many-type stream (Zipf over >= 2 M synthetic types) with a 32 K model."""
    import torch
    from subword_tokenizers_b200 import device, packing as P
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    mat, lens = cached_synth_table(args.many_types, 1, rank, dist, world)
    t_arena, t_off = BD.table_to_utf8(mat, lens)
    zs = BD.ZipfStream(t_arena, t_off, ZIPF_C, 100 + rank)
    d_arena, d_off, n_words, _ = zs.device_stream(args.many_bytes, dev)
    n_bytes = int(d_arena.numel())
    pre = BD.ZipfStream(t_arena, t_off, ZIPF_C, 100 + rank).host_sample(min(args.check_words // 4, n_words))
    out = {"workload": "BASELINE configs[3] shape: %.2f GB per GPU, i.i.d. Zipf(C=%d) stream over %d synthetic word types (the tail beyond "
                       "rank C has weight 1), 32 K model: the %d merges of the 10 M-type training above (FastBPE) and a vocabulary of "
                       "their symbols (FastWP)" % (n_bytes / 1e9, ZIPF_C, args.many_types, len(merges)),
           "n_words": n_words, "n_bytes": n_bytes, "n_types": int(args.many_types)}
    # 32 K WordPiece vocabulary derived from the trained symbols: every character, "##" + character, and the first symbols as
    # word-initial and as continuation entries
    half = max(0, (32000 - 2 * n_alpha) // 2)
    vocab = set(strs[:n_alpha]) | {"##" + s for s in strs[:n_alpha]} | set(strs[n_alpha:n_alpha + half]) | {"##" + s for s in strs[n_alpha:n_alpha + half]}
    threads = os.cpu_count() or 1
    for name, kind in (("fastwp", "wp"), ("fastbpe", "bpe")):
        if kind == "wp":
            tab = P.WpTables(vocab)
            enc = device.WpEncoder(tab, naive_wp_encode_ids("##", tab))
        else:
            tab = P.BpeTables(merges)
            enc = device.BpeEncoder(tab)
        r = time_encode(enc, d_arena, d_off, n_words, max(3, min(args.steps, 5)), 2, dist, world)
        max_ms, job_bytes = reduce_max_sum(dist, world, dev, r["total_ms"], n_bytes)
        steps = len(r["step_ms"])
        alg = n_bytes + 8 * (n_words + 1) + 4 * r["n_tokens"]
        npre, ok, _, _ = check_prefix(kind, tab, pre, r["d_ids"], r["d_tok"], threads)
        out[name] = {"value": job_bytes * steps / (max_ms / 1e3) / 1e6, "unit": "MB/s", "ms_per_step": max_ms / steps, "n_tokens": r["n_tokens"],
                     "model_entries": tab.n_vocab if kind == "wp" else len(merges), "memo_types": r["memo_types"], "slow_words": r["slow_words"],
                     "parity_checked_words": npre, "parity_ok": ok,
                     "roofline": roofline(alg, float(np.mean(r["step_ms"])), "one %s encode call (7 kernels)" % name)}
        del r
        enc.close()
    return out


def small_trainings(dev):
    """train-5K (22,971 types) at max_vocab 8000: BPE and WordPiece merge loops on one GPU, compared with the reference's lists."""
    import torch
    from subword_tokenizers_b200 import device, packing as P, _lib as L
    from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
    out = {}
    merges = [tuple(p) for p in load_golden("ref_bpe_train5k_v8000_merges.json.gz")]
    pre = make_hf_tokenizer().backend_tokenizer.pre_tokenizer
    words = [w for s in load_golden("train-5K.json.gz") for w, _ in pre.pre_tokenize_str(s.lower())]
    tt = P.TrainTypes(words)
    max_len = int(np.diff(tt.off.astype(np.int64)).max())
    best = None
    for _ in range(2):
        eng = device.CudaTrainEngine(tt.syms, tt.off, tt.freq, tt.n_alpha, 8000, tt.n_alpha, max_len, 0, 0, 1, record_cap=8192)
        torch.cuda.synchronize()
        tic = time.perf_counter()
        l, r, n, c, state = device.run_training_loop(eng, 1, steps_per_sync=1024)
        torch.cuda.synchronize()
        dt = time.perf_counter() - tic
        best = dt if best is None else min(best, dt)
        eng.close()
    m, _ = tt.merges_to_strs(l, r, n)
    out["bpe_train"] = {"value": len(l) / best, "unit": "merges/s", "merges": int(len(l)), "seconds": best,
                        "n_types": tt.n_types, "n_symbols": int(len(tt.syms)), "matches_reference_merges": m == merges,
                        "workload": "train-5K word types, max_vocab 8000 (merge loop only)"}
    wt = P.WpTrainTypes(words)
    wmax = int(np.diff(wt.off.astype(np.int64)).max())
    best = None
    for _ in range(2):
        eng = device.CudaTrainEngine(wt.syms, wt.off, wt.freq, len(wt.init_syms), 8000, len(wt.init_syms), wmax + 2, 0, 0, 1,
                                     record_cap=8192, mode=L.TRAIN_WP, init_cps=wt.init_cps, init_off=wt.init_off)
        torch.cuda.synchronize()
        tic = time.perf_counter()
        l, r, n, c, state = device.run_training_loop(eng, 1, steps_per_sync=1024)
        torch.cuda.synchronize()
        dt = time.perf_counter() - tic
        best = dt if best is None else min(best, dt)
        eng.close()
    vocab = sorted(wt.vocab_from_merges(l, r, n))
    out["wp_train"] = {"value": len(l) / best, "unit": "merges/s", "merges": int(len(l)), "seconds": best,
                       "matches_reference_vocab": vocab == load_golden("ref_wp_train5k_v8000_vocab.json.gz"),
                       "workload": "NaiveWP.train, train-5K word types, max_vocab 8000 (merge loop only)"}
    return out


def pretok_numbers(dev, d_arena, d_off, n_words, n_bytes):
    """Device pre-tokenization (lower-casing + whitespace split, FastWP) over the raw-text form of the stream, resident."""
    import torch
    from subword_tokenizers_b200 import device
    lib = device._lib.load()
    d_text, n_text = device_text(d_arena, d_off, n_words)
    pt = device.Pretokenizer.get()
    pws = torch.empty(lib.swt_pretok_workspace_bytes(n_text), dtype=torch.uint8, device=dev)
    o_arena = torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev)
    o_off = torch.empty(n_words + 2, dtype=torch.int32, device=dev)
    status = torch.empty(8, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_cnt = t_wr = 0.0
    for it in range(4):
        ev[0].record()
        device.check(lib.swt_pretok_count(pt._handle, d_text.data_ptr(), n_text, pws.data_ptr(), pws.numel(), status.data_ptr(), sp))
        ev[1].record()
        device.check(lib.swt_pretok_write(pt._handle, d_text.data_ptr(), n_text, pws.data_ptr(), pws.numel(), o_arena.data_ptr(), n_bytes,
                                          o_off.data_ptr(), None, n_words + 2, n_words, n_bytes, status.data_ptr(), sp))
        ev[2].record(); torch.cuda.synchronize()
        if it:
            t_cnt += ev[0].elapsed_time(ev[1]) / 3; t_wr += ev[1].elapsed_time(ev[2]) / 3
    same = bool(torch.equal(o_arena[:n_bytes], d_arena[:n_bytes]) and torch.equal(o_off[:n_words + 1], d_off[:n_words + 1]))
    # algorithmic bytes: the text is read by both passes, the arena and 4 B per word are written
    alg_w = n_text + n_bytes + 4 * (n_words + 1)
    return {"value": n_text / ((t_cnt + t_wr) / 1e3) / 1e6, "unit": "MB/s of raw text", "text_bytes": n_text,
            "count_ms": t_cnt, "write_ms": t_wr, "reproduces_the_packed_stream": same,
            "roofline_count": roofline(n_text, t_cnt, "pretok_kernel<count>"), "roofline": roofline(alg_w, t_wr, "pretok_kernel<write>"),
            "workload": "text.lower().split() on the device over the stream's words joined by spaces"}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=1_000_000_000, help="stream bytes per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-also", action="store_true", help="skip the secondary numbers")
    ap.add_argument("--train-types", type=int, default=10_000_000)
    ap.add_argument("--train-vocab", type=int, default=32_000)
    ap.add_argument("--check-oracle-steps", type=int, default=3)
    ap.add_argument("--many-types", type=int, default=2_000_000)
    ap.add_argument("--many-bytes", type=int, default=500_000_000)
    ap.add_argument("--check-words", type=int, default=4_000_000, help="words of the stream prefix compared with the oracle at N > 1")
    ap.add_argument("--e2e-batch-mb", type=int, default=32, help="batch size of the host-buffer pipeline (MiB of input per batch)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vocab = load_golden("ref_wp_train5k_v8000_vocab.json.gz")
    workload = ("FastWP tokenize, vocab 8000 trained by the reference on train-5K (train-85k absent), "
                "%.2f GB synthetic Zipf(s=1, C=%d) stream of the 22,971 train-5K word types per GPU" % (args.bytes / 1e9, ZIPF_C))
    config = {"workload": workload, "bytes_per_gpu": args.bytes, "seed": args.seed,
              "l2": "input (1 GB) and output exceed the 126 MB L2: no flush between iterations"}
    threads = os.cpu_count() or 1

    # ---------------------------------------------------------------- reference arm: CPU port, all host threads
    if args.impl == "reference":
        if rank != 0:
            return
        from subword_tokenizers_b200 import packing as P
        stream = ZipfStream.train5k(args.seed)
        sample_words = min(int(args.bytes / stream.mean_len), 4_000_000 * max(1, threads // 2))
        arena, off = stream.host_sample(sample_words)
        tab = P.WpTables(vocab)
        oracle_encode("wp", tab, arena[: int(off[200_000])], off[:200_001], threads)          # warm-up
        best = None
        for _ in range(max(1, args.steps)):
            _, sec = oracle_encode("wp", tab, arena, off, threads)
            best = sec if best is None else min(best, sec)
        mbps = len(arena) / best / 1e6
        config["reference_sample"] = "a PREFIX of the workload: the first %d words (%.1f MB) of the rank-0 stream per step" % (sample_words, len(arena) / 1e6)
        line = {"impl": "reference", "metric": "FastWP tokenize MB/s", "value": mbps, "unit": "MB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": best * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": mbps, "unit": "MB/s", "cores": threads, "kind": "port",
                                 "sample": "first %d words (%.1f MB) of the rank-0 stream, best of %d" % (sample_words, len(arena) / 1e6, max(1, args.steps))},
                "e2e": {"value": mbps, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from subword_tokenizers_b200 import device, packing as P
    from subword_tokenizers_b200.utils import naive_wp_encode_ids
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = ZipfStream.train5k(args.seed + rank)                  # weak scaling: every rank encodes its own stream
    d_arena, d_off, n_words, off32 = stream.device_stream(args.bytes, dev)
    n_bytes = int(d_arena.numel())
    tab = P.WpTables(vocab)
    enc = device.WpEncoder(tab, naive_wp_encode_ids("##", tab))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    r = time_encode(enc, d_arena, d_off, n_words, args.steps, args.warmup, dist, world)
    clocks = sampler.summary() if rank == 0 else None
    n_tokens, h6 = r["n_tokens"], r["h6"]
    alg_bytes = n_bytes + 4 * (n_words + 1) + 4 * n_tokens + 4 * (n_words + 1)
    max_ms, job_bytes = reduce_max_sum(dist, world, dev, r["total_ms"], n_bytes)
    value = job_bytes * args.steps / (max_ms / 1e3) / 1e6
    kern_ms = float(np.mean(r["step_ms"]))
    # ---- parity of THIS run's output with the CPU oracle on a prefix of the stream (every rank checks its own stream)
    pre_words = min(n_words, 4_000_000 * max(1, threads) if world == 1 else args.check_words)
    prefix = ZipfStream.train5k(args.seed + rank).host_sample(pre_words)
    npre, parity_ok, oracle_sec, pre_bytes = check_prefix("wp", tab, prefix, r["d_ids"], r["d_tok"], threads)
    pt = torch.tensor([1.0 if parity_ok else 0.0, float(npre)], dtype=torch.float64, device=dev)
    if world > 1:
        mn = pt.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
        parity_all, parity_words = bool(mn[0].item() == 1.0), int(pt[1].item())
    else:
        parity_all, parity_words = parity_ok, npre
    ref_ids = r["d_ids"][:n_tokens].cpu()
    d_ids = r["d_ids"]
    del r

    # ---- end to end through the host-buffer C ABI (pinned host buffers, copies inside the timed region)
    h_arena = torch.empty(n_bytes, dtype=torch.uint8).pin_memory(); h_arena.copy_(d_arena)
    h_off = torch.from_numpy(off32.view(np.int32)).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def time_e2e(h_ids, h_tok):
        enc.encode_host(h_arena, h_off, h_ids, h_tok)          # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tic = time.perf_counter()
        for _ in range(e2e_steps):
            nt, _ = enc.encode_host(h_arena, h_off, h_ids, h_tok)
        torch.cuda.synchronize()
        sec = time.perf_counter() - tic
        assert nt == n_tokens
        t2 = torch.tensor([sec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        return job_bytes * e2e_steps / float(t2.item()) / 1e6

    h_ids16 = torch.empty(n_tokens + 1024, dtype=torch.int16).pin_memory()
    e2e_value = time_e2e(h_ids16, None)
    e2e_ok = bool(torch.equal(h_ids16[:n_tokens].to(torch.int32) & 0xFFFF, ref_ids))
    # raw text (words joined by single spaces) -> device pre-tokenization (lower + whitespace split) + encode -> 16-bit ids
    d_text, n_text = device_text(d_arena, d_off, n_words)
    h_text = torch.empty(d_text.numel(), dtype=torch.uint8).pin_memory(); h_text.copy_(d_text)
    del d_text
    h_ids16.zero_()
    e2e_batch = args.e2e_batch_mb << 20
    enc.tokenize_host(h_text, n_text, h_ids16, has_sigma=False, batch_bytes=e2e_batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tic = time.perf_counter()
    for _ in range(e2e_steps):
        nt_text, nw_text, _ = enc.tokenize_host(h_text, n_text, h_ids16, has_sigma=False, batch_bytes=e2e_batch)
    torch.cuda.synchronize()
    t3 = torch.tensor([time.perf_counter() - tic], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    assert nt_text == n_tokens and nw_text == n_words
    e2e_text_sec = float(t3.item()) / e2e_steps
    e2e_text_value = job_bytes / e2e_text_sec / 1e6
    e2e_text_ok = bool(torch.equal(h_ids16[:n_tokens].to(torch.int32) & 0xFFFF, ref_ids))
    del h_ids16, h_text
    h_ids = torch.empty(n_tokens + 1024, dtype=torch.int32).pin_memory()
    h_tok = torch.empty(n_words + 1, dtype=torch.int32).pin_memory()
    e2e32_value = time_e2e(h_ids, h_tok)
    e2e32_ok = bool(torch.equal(h_ids[:n_tokens], ref_ids))
    del h_ids, h_tok, h_arena, h_off, ref_ids
    # the same copies without kernels: the PCIe ceiling of the headline e2e number (all ranks concurrently)
    ceil_sec = copy_ceiling(dev, n_text, 2 * n_tokens, e2e_batch, e2e_steps, dist, world)

    line = None
    if rank == 0:
        line = {
            "metric": "FastWP tokenize MB/s", "value": value, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
            "roofline": roofline(alg_bytes, kern_ms, "memo_clear + encode_count_kernel<WpEnc> + encode_long_count + 2 scans + "
                                 "encode_emit_kernel<WpEnc> + encode_long_emit (one encode call)", "fastwp_1GB"),
            "parity": {"checked_against": "oracle (C port of reference wordpiece.py:233-316) on the first words of each rank's stream",
                       "parity_checked_words": parity_words, "ok": parity_all},
            "e2e": {"value": e2e_text_value, "unit": "MB/s", "h2d_bytes_per_step": n_text,
                    "d2h_bytes_per_step": 2 * n_tokens + 64 * (n_text // e2e_batch + 1), "steps": e2e_steps, "batch_mib": args.e2e_batch_mb,
                    "call": "swt_tokenize_text_host",
                    "what": "raw UTF-8 text (the stream's words joined by single spaces) in a pinned host buffer -> H2D -> lower-casing + "
                            "whitespace split on the device (swt_pretok_*) -> FastWP encode -> 16-bit flat token ids (the list "
                            "tokenize() returns) D2H into a pinned host buffer; MB = word bytes, as in `value`",
                    "matches_resident_run": e2e_text_ok,
                    "copy_ceiling": {"seconds_per_step": ceil_sec, "value": job_bytes / ceil_sec / 1e6, "unit": "MB/s",
                                     "what": "the same H2D + D2H byte counts through five streams in %d MiB batches, no kernels, " % args.e2e_batch_mb +
                                             "all %d rank(s) concurrently (max over ranks)" % world},
                    "frac_of_copy_ceiling": ceil_sec / e2e_text_sec,
                    "packed_words_in_16bit_ids": {"value": e2e_value, "unit": "MB/s", "call": "swt_encode_host16",
                                                  "h2d_bytes_per_step": n_bytes + 4 * (n_words + 1),
                                                  "d2h_bytes_per_step": 2 * n_tokens + 32 * ((n_bytes >> 26) + 1), "matches_resident_run": e2e_ok},
                    "packed_words_in_u32_ids_and_word_offsets": {"value": e2e32_value, "unit": "MB/s", "call": "swt_encode_host",
                                                                 "h2d_bytes_per_step": n_bytes + 4 * (n_words + 1),
                                                                 "d2h_bytes_per_step": 4 * n_tokens + 4 * n_words + 32 * ((n_bytes >> 26) + 1),
                                                                 "matches_resident_run": e2e32_ok}},
            "gpu_launches": KERNELS_PER_ENCODE_CALL * args.steps, "clocks": clocks,
            "stream": {"n_words": n_words, "n_bytes": n_bytes, "n_tokens": n_tokens, "h6_events": h6},
        }
        if world == 1:
            # the oracle run of the parity check IS the CPU baseline: all host threads over the prefix
            line["cpu_baseline"] = {"value": pre_bytes / oracle_sec / 1e6, "unit": "MB/s", "cores": threads, "kind": "port",
                                    "sample": "first %d words (%.1f MB) of the stream, %.2f s; its ids are the ones the GPU output "
                                              "was compared with" % (npre, pre_bytes / 1e6, oracle_sec)}
            pyref = python_reference_rate(vocab, prefix[0], prefix[1])
            if pyref is not None:
                line["cpu_baseline"]["python_reference_one_core"] = pyref
    del d_ids
    also = {}
    if not args.no_also:
        # ---- FastBPE over the same stream (7,922 merges trained by the reference on train-5K), every N
        merges5k = [tuple(p) for p in load_golden("ref_bpe_train5k_v8000_merges.json.gz")]
        btab = P.BpeTables(merges5k)
        benc = device.BpeEncoder(btab)
        rb = time_encode(benc, d_arena, d_off, n_words, max(3, min(args.steps, 10)), 2, dist, world)
        bsteps = len(rb["step_ms"])
        bmax_ms, _ = reduce_max_sum(dist, world, dev, rb["total_ms"], n_bytes)
        bpre = (prefix[0][: int(prefix[1][min(npre, args.check_words)])], prefix[1][: min(npre, args.check_words) + 1])
        bn, bok, _, _ = check_prefix("bpe", btab, bpre, rb["d_ids"], rb["d_tok"], threads)
        balg = n_bytes + 8 * (n_words + 1) + 4 * rb["n_tokens"]
        also["fastbpe_tokenize"] = {"value": job_bytes * bsteps / (bmax_ms / 1e3) / 1e6, "unit": "MB/s", "merges": len(merges5k),
                                    "ms_per_step": bmax_ms / bsteps, "n_tokens": rb["n_tokens"], "parity_checked_words": bn, "parity_ok": bok,
                                    "roofline": roofline(balg, float(np.mean(rb["step_ms"])), "one FastBPE encode call (7 kernels)", "fastbpe_1GB")}
        del rb
        benc.close()
        if world == 1:
            also["fastwp_pretokenize"] = pretok_numbers(dev, d_arena, d_off, n_words, n_bytes)
            # direct-path rates: the word-type memo disabled, every word is encoded (trie walk / merge loop) -- first 200 MB of the stream
            try:
                nw_d = int(torch.searchsorted(d_off[: n_words + 1].long(), 200_000_000).item())
                nb_d = int(d_off[nw_d].item())
                device.tune("memo_off", 1)
                for name, e_d, kind, t_d in (("fastwp_direct", enc, "wp", tab), ("fastbpe_direct", None, "bpe", None)):
                    if e_d is None:
                        e_d, t_d = device.BpeEncoder(btab), btab
                    rd = time_encode(e_d, d_arena[:nb_d + 64], d_off[: nw_d + 1], nw_d, 3, 2, dist, 1)
                    dn, dok, _, _ = check_prefix(kind, t_d, (prefix[0][: int(prefix[1][200_000])], prefix[1][:200_001]), rd["d_ids"], rd["d_tok"], threads)
                    also[name] = {"value": nb_d * len(rd["step_ms"]) / (rd["total_ms"] / 1e3) / 1e6, "unit": "MB/s", "ms_per_step": float(np.mean(rd["step_ms"])),
                                  "bytes": nb_d, "parity_checked_words": dn, "parity_ok": dok,
                                  "what": "swt_tune(\"memo_off\", 1): no word-type memo, every word goes through the batched direct path"}
                    del rd
                    if kind == "bpe":
                        e_d.close()
            except Exception as e:                              # noqa: BLE001
                also["direct_error"] = repr(e)
            finally:
                device.tune("memo_off", 0)
    del d_arena, d_off
    enc.close()
    torch.cuda.empty_cache()
    if not args.no_also:
        try:
            tr, merges, strs, n_alpha = train_10m(args, dev, rank, world, dist)
            also["bpe_train_10M"] = tr
            also["many_types"] = many_types(args, dev, rank, world, dist, merges, strs, n_alpha)
        except Exception as e:                                  # noqa: BLE001 -- the headline line must still be printed
            also["error"] = repr(e)
        if world == 1:
            try:
                also.update(small_trainings(dev))
            except Exception as e:                              # noqa: BLE001
                also["small_trainings_error"] = repr(e)
    if rank == 0:
        if also:
            line["also"] = also
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
