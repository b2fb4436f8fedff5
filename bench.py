#!/usr/bin/env python
"""bench.py -- headline benchmark of the tokenize/train hot paths (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bytes B]

Workload (BASELINE.json configs[1]): FastWP, vocabulary of 8000 trained by the UNMODIFIED reference on
data/train-5K.json (data/train-85k.json is missing from the reference checkout; fixture
tests/golden/ref_wp_train5k_v8000_vocab.json.gz), tokenizing a 1 GB synthetic Zipf word stream per GPU.
A "step" is one pass of the FastWP encode kernel over the whole stream.

  value      MB/s (10^6 input arena bytes per second) with the stream resident in HBM, CUDA-event timed,
             max over ranks, whole job.
  e2e        same metric through the host-buffer C ABI from RAW TEXT (swt_tokenize_text_host): pinned host text in,
             lower-casing + whitespace split + encode on the device, flat 16-bit token ids out; H2D and D2H inside the
             timed region.  Sub-entries: the packed-words variants (swt_encode_host16 / swt_encode_host).
  roofline   algorithmic bytes (arena + 4 B/word offset read, 4 B/token + 4 B/word offset written) per launch
             / mean kernel time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the C oracle port of the reference's FastWP path on the host cores, bounded sample.
  also       secondary numbers: FastBPE tokenize MB/s, device pre-tokenization MB/s, BPE / WordPiece train merges/s.

--impl reference: times the CPU implementation of the same path (oracle C port -- the reference itself is
pure Python and cannot travel to the GPU box) with all host threads on a bounded sample of the same stream.
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt", encoding="utf-8") as f:
        return json.load(f)


# ------------------------------------------------------------------------------------------------------------
# synthetic Zipf stream: types = the word types of train-5K (BERT pre-tokenized, lower-cased),
# ranked by a seeded permutation; rank r has integer weight max(1, floor(C / r)); words are i.i.d. draws.
# The draw sequence comes from numpy PCG64 in fixed chunks, so any prefix is reproducible on the CPU.
# ------------------------------------------------------------------------------------------------------------
ZIPF_C = 200_000
# dram__bytes_read.sum + dram__bytes_write.sum of the kernels of ONE FastWP encode call over the 1 GB bench stream
# (ncu --set full, see profiles/r01_final_ncu_wp_1GB.txt); None until measured
TRAFFIC_1GB_WP = 6_195_873_000
TRAFFIC_SOURCE = "profiles/r01_final6_ncu_wp_1GB.txt"
DRAW_CHUNK = 1 << 24


class ZipfStream:
    def __init__(self, seed: int):
        from subword_tokenizers_b200 import packing as P
        from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
        pre = make_hf_tokenizer().backend_tokenizer.pre_tokenizer
        corpus = load_golden("train-5K.json.gz")
        types = list(dict.fromkeys(w for s in corpus for w, _ in pre.pre_tokenize_str(s.lower())))
        rng = np.random.Generator(np.random.PCG64(seed))
        order = rng.permutation(len(types))
        self.types = [types[i] for i in order]
        weights = np.maximum(1, ZIPF_C // np.arange(1, len(types) + 1)).astype(np.int64)
        self.lut = np.repeat(np.arange(len(types), dtype=np.int32), weights)
        self.t_arena, t_off = P.pack_words(self.types)
        self.t_off = t_off.astype(np.int64)
        self.t_len = np.diff(self.t_off)
        self.mean_len = float((self.t_len[self.lut]).mean())
        self.rng = np.random.Generator(np.random.PCG64(seed + 1000))
        self.seed = seed

    def draws(self, n_words: int):
        """Yields int32 chunks of type ids (deterministic sequence)."""
        left = n_words
        while left > 0:
            k = min(left, DRAW_CHUNK)
            full = self.rng.integers(0, len(self.lut), size=DRAW_CHUNK, dtype=np.int64)   # always a full chunk: prefix-stable
            yield self.lut[full[:k]]
            left -= k

    def host_sample(self, n_words: int):
        """First n_words of the stream as (arena u8, offsets u64) on the host."""
        draw = np.concatenate(list(self.draws(n_words)))
        lens = self.t_len[draw]
        off = np.zeros(n_words + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        idx = np.repeat(self.t_off[:-1][draw] - off[:-1], lens) + np.arange(off[-1])
        return self.t_arena[idx], off.astype(np.uint64)

    def device_stream(self, target_bytes: int, dev):
        """The stream resident on `dev`: (arena u8 tensor, offsets int32-view tensor (u32), n_words)."""
        import torch
        n_words = int(target_bytes / self.mean_len)
        d_tlen = torch.from_numpy(self.t_len).to(dev)
        d_toff = torch.from_numpy(self.t_off[:-1].copy()).to(dev)
        d_tarena = torch.from_numpy(self.t_arena).to(dev)
        d_draw = torch.empty(n_words, dtype=torch.int32, device=dev)
        p = 0
        for chunk in self.draws(n_words):
            d_draw[p:p + len(chunk)] = torch.from_numpy(chunk).to(dev)
            p += len(chunk)
        lens = d_tlen[d_draw.long()]
        off = torch.zeros(n_words + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=off[1:])
        total = int(off[-1].item())
        assert total < (1 << 32) - 64
        arena = torch.empty(total, dtype=torch.uint8, device=dev)
        step = 1 << 23
        for a in range(0, n_words, step):
            b = min(n_words, a + step)
            l = lens[a:b]
            o = off[a:b] - off[a]
            nb = int((off[b] - off[a]).item())
            widx = torch.repeat_interleave(torch.arange(b - a, device=dev), l, output_size=nb)
            src = d_toff[d_draw[a:b].long()][widx] + (torch.arange(nb, device=dev) - o[widx])
            arena[int(off[a].item()):int(off[b].item())] = d_tarena[src]
        off32 = off.to(torch.int64).cpu().numpy().astype(np.uint32)
        d_off = torch.from_numpy(off32.view(np.int32)).to(dev)
        return arena, d_off, n_words, off32


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


def cpu_oracle_wp(stream: "ZipfStream", vocab, sample_words: int, threads: int, repeats: int = 1):
    """Times the C oracle port of FastWP.tokenize on the first sample_words words, `threads` host threads
    (ctypes releases the GIL; every thread owns a disjoint slice). -> (MB/s, seconds, bytes)."""
    import oracle
    from subword_tokenizers_b200 import packing as P
    arena, off = stream.host_sample(sample_words)
    tab = P.WpTables(vocab)
    alnum, space = P.unicode_class_bitmaps()
    tries = [oracle.WpTrie(tab, alnum) for _ in range(threads)]
    bounds = np.linspace(0, sample_words, threads + 1).astype(np.int64)
    slices = []
    for t in range(threads):
        a, b = int(bounds[t]), int(bounds[t + 1])
        o = off[a:b + 1] - off[a]
        slices.append((arena[int(off[a]):int(off[b])].copy(), o.copy()))

    def work(t):
        ids, _, _ = tries[t].encode(slices[t][0], slices[t][1], space)
        return len(ids)
    best = None
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(repeats):
            tic = time.perf_counter()
            n_tok = sum(ex.map(work, range(threads)))
            dt = time.perf_counter() - tic
            best = dt if best is None else min(best, dt)
    return len(arena) / best / 1e6, best, len(arena), n_tok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=1_000_000_000, help="stream bytes per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-also", action="store_true", help="skip the secondary FastBPE / training numbers")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vocab = load_golden("ref_wp_train5k_v8000_vocab.json.gz")
    workload = ("FastWP tokenize, vocab 8000 trained by the reference on train-5K (train-85k absent), "
                "%.2f GB synthetic Zipf(s=1, C=%d) stream of the 22,971 train-5K word types per GPU" % (args.bytes / 1e9, ZIPF_C))
    config = {"workload": workload, "bytes_per_gpu": args.bytes, "seed": args.seed,
              "l2": "input (1 GB) and output exceed the 126 MB L2: no flush between iterations"}

    # ---------------------------------------------------------------- reference arm: CPU port, all host threads
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        stream = ZipfStream(args.seed)
        sample_words = min(int(args.bytes / stream.mean_len), 4_000_000 * max(1, threads // 2))
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_oracle_wp(stream.__class__(args.seed), vocab, min(sample_words, 200_000), threads)
        times, nbytes = [], 0
        stream = ZipfStream(args.seed)
        mbps, sec, nbytes, _ = cpu_oracle_wp(stream, vocab, sample_words, threads, repeats=max(1, args.steps))
        line = {"impl": "reference", "metric": "FastWP tokenize MB/s", "value": mbps, "unit": "MB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": mbps, "unit": "MB/s", "cores": threads, "kind": "port",
                                 "sample": "first %d words (%.1f MB) of the rank-0 stream, best of %d" % (sample_words, nbytes / 1e6, max(1, args.steps))},
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
    stream = ZipfStream(args.seed + rank)                  # weak scaling: every rank encodes its own stream
    d_arena, d_off, n_words, off32 = stream.device_stream(args.bytes, dev)
    n_bytes = int(d_arena.numel())
    tab = P.WpTables(vocab)
    enc = device.WpEncoder(tab, naive_wp_encode_ids("##", tab))
    lib_ws = device._lib.load().swt_encode_workspace_bytes(n_words, 0)
    d_ws = torch.empty(lib_ws, dtype=torch.uint8, device=dev)
    out_cap = n_bytes + n_words + 16
    d_ids = torch.empty(out_cap, dtype=torch.int32, device=dev)
    d_tok = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
    d_status = torch.empty(8, dtype=torch.int32, device=dev)

    def one_pass():
        enc.encode_into(d_arena, d_off, n_words, 0, d_ids, out_cap, d_tok, d_ws, d_status)

    for _ in range(max(3, args.warmup)):
        one_pass()
    n_tokens, h6 = enc.check_status(d_status)
    alg_bytes = n_bytes + 4 * (n_words + 1) + 4 * n_tokens + 4 * (n_words + 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(args.steps):
        one_pass()
        ev[k + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    clocks = sampler.summary() if rank == 0 else None
    enc.check_status(d_status)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    tot_bytes = torch.tensor([float(n_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_bytes, op=dist.ReduceOp.SUM)
    max_ms, job_bytes = float(t.item()), float(tot_bytes.item())
    value = job_bytes * args.steps / (max_ms / 1e3) / 1e6

    # ---- end to end through the host-buffer C ABI (pinned host buffers, copies inside the timed region).
    # Headline: swt_encode_host16 returning the flat token list the reference's tokenize() returns (16-bit ids, the
    # vocabulary has 8002 ids).  Also timed: the 32-bit ids + per-word token offsets variant (swt_encode_host).
    h_arena = torch.empty(n_bytes, dtype=torch.uint8).pin_memory(); h_arena.copy_(d_arena)
    h_off = torch.from_numpy(off32.view(np.int32)).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    ref_ids = d_ids[:n_tokens].cpu()

    def time_e2e(h_ids, h_tok):
        enc.encode_host(h_arena, h_off, h_ids, h_tok)          # warm-up (allocates the pipeline slots)
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
    enc.tokenize_host(h_text, n_text, h_ids16, has_sigma=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tic = time.perf_counter()
    for _ in range(e2e_steps):
        nt_text, nw_text, _ = enc.tokenize_host(h_text, n_text, h_ids16, has_sigma=False)
    torch.cuda.synchronize()
    t3 = torch.tensor([time.perf_counter() - tic], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    assert nt_text == n_tokens and nw_text == n_words
    e2e_text_value = job_bytes * e2e_steps / float(t3.item()) / 1e6
    e2e_text_ok = bool(torch.equal(h_ids16[:n_tokens].to(torch.int32) & 0xFFFF, ref_ids))
    del h_ids16, h_text
    h_ids = torch.empty(n_tokens + 1024, dtype=torch.int32).pin_memory()
    h_tok = torch.empty(n_words + 1, dtype=torch.int32).pin_memory()
    e2e32_value = time_e2e(h_ids, h_tok)
    e2e32_ok = bool(torch.equal(h_ids[:n_tokens], ref_ids))
    del h_ids, h_tok

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    kern_ms = float(np.mean(step_ms))
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    line = {
        "metric": "FastWP tokenize MB/s", "value": value, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": TRAFFIC_1GB_WP if args.bytes == 1_000_000_000 else None, "traffic_source": TRAFFIC_SOURCE,
                     "peak_source": peak_src, "kernel": "encode_count_kernel<WpEnc> + scan + encode_emit_kernel<WpEnc> (one encode call)",
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms},
        "e2e": {"value": e2e_text_value, "unit": "MB/s", "h2d_bytes_per_step": n_text,
                "d2h_bytes_per_step": 2 * n_tokens + 64 * ((n_text >> 26) + 1), "steps": e2e_steps,
                "call": "swt_tokenize_text_host",
                "what": "raw UTF-8 text (the stream's words joined by single spaces) in a pinned host buffer -> H2D -> lower-casing + "
                        "whitespace split on the device (swt_pretok_*) -> FastWP encode -> 16-bit flat token ids (the list "
                        "tokenize() returns) D2H into a pinned host buffer; MB = word bytes, as in `value`",
                "matches_resident_run": e2e_text_ok,
                "packed_words_in_16bit_ids": {"value": e2e_value, "unit": "MB/s", "call": "swt_encode_host16",
                                              "h2d_bytes_per_step": n_bytes + 4 * (n_words + 1),
                                              "d2h_bytes_per_step": 2 * n_tokens + 32 * ((n_bytes >> 26) + 1), "matches_resident_run": e2e_ok},
                "packed_words_in_u32_ids_and_word_offsets": {"value": e2e32_value, "unit": "MB/s", "call": "swt_encode_host",
                                                             "h2d_bytes_per_step": n_bytes + 4 * (n_words + 1),
                                                             "d2h_bytes_per_step": 4 * n_tokens + 4 * n_words + 32 * ((n_bytes >> 26) + 1),
                                                             "matches_resident_run": e2e32_ok}},
        "gpu_launches": 5 * args.steps, "clocks": clocks,
        "stream": {"n_words": n_words, "n_bytes": n_bytes, "n_tokens": n_tokens, "h6_events": h6},
    }
    # ---- CPU baseline: oracle port on the host cores, bounded sample (N=1 only)
    if args.gpus == 1:
        threads = os.cpu_count() or 1
        sample_words = min(n_words, 4_000_000 * max(1, threads))
        mbps, sec, nb, _ = cpu_oracle_wp(ZipfStream(args.seed), vocab, sample_words, threads)
        line["cpu_baseline"] = {"value": mbps, "unit": "MB/s", "cores": threads, "kind": "port",
                                "sample": "first %d words (%.1f MB) of the stream, %.2f s" % (sample_words, nb / 1e6, sec)}
        if not args.no_also:
            line["also"] = secondary_numbers(dev, stream, d_arena, d_off, n_words, n_bytes)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def secondary_numbers(dev, stream, d_arena, d_off, n_words, n_bytes):
    """FastBPE tokenize (same stream, 7,922 merges trained by the reference on train-5K) and BPE training
    (train-5K word types, max_vocab 8000) on one GPU; reported beside the headline, not as it."""
    import torch
    from subword_tokenizers_b200 import device, packing as P
    out = {}
    merges = [tuple(p) for p in load_golden("ref_bpe_train5k_v8000_merges.json.gz")]
    enc = device.BpeEncoder(P.BpeTables(merges))
    ws = torch.empty(device._lib.load().swt_encode_workspace_bytes(n_words, 0), dtype=torch.uint8, device=dev)
    cap = n_bytes + n_words + 16
    ids = torch.empty(cap, dtype=torch.int32, device=dev)
    tok = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
    status = torch.empty(8, dtype=torch.int32, device=dev)
    for _ in range(2):
        enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        enc.encode_into(d_arena, d_off, n_words, 0, ids, cap, tok, ws, status)
    b.record(); torch.cuda.synchronize()
    nt, _ = enc.check_status(status)
    ms = a.elapsed_time(b) / 3
    peak, _ = hbm_peak()
    alg = n_bytes + 8 * (n_words + 1) + 4 * nt
    out["fastbpe_tokenize"] = {"value": n_bytes / (ms / 1e3) / 1e6, "unit": "MB/s", "merges": len(merges), "kernel_ms": ms,
                               "n_tokens": nt, "roofline_frac": alg / (ms / 1e3) / 1e9 / peak}
    del ids, tok, ws
    # device pre-tokenization (lower-casing + whitespace split, FastWP) over the raw-text form of the stream, resident
    lib = device._lib.load()
    d_text, n_text = device_text(d_arena, d_off, n_words)
    pt = device.Pretokenizer.get()
    pws = torch.empty(lib.swt_pretok_workspace_bytes(n_text), dtype=torch.uint8, device=dev)
    o_arena = torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev)
    o_off = torch.empty(n_words + 2, dtype=torch.int32, device=dev)
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
    out["fastwp_pretokenize"] = {"value": n_text / ((t_cnt + t_wr) / 1e3) / 1e6, "unit": "MB/s of raw text", "text_bytes": n_text,
                                 "count_ms": t_cnt, "write_ms": t_wr, "reproduces_the_packed_stream": same,
                                 "workload": "text.lower().split() on the device over the stream's words joined by spaces"}
    del d_text, pws, o_arena, o_off
    # BPE training, config-1 corpus at max_vocab 8000
    from subword_tokenizers_b200.hf_shim import make_hf_tokenizer
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
                        "n_types": tt.n_types, "n_symbols": int(len(tt.syms)),
                        "matches_reference_merges": m == merges,
                        "workload": "train-5K word types, max_vocab 8000 (merge loop only)"}
    # NaiveWP.train (next row of the scope table), same corpus, max_vocab 8000
    from subword_tokenizers_b200 import _lib as L
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


if __name__ == "__main__":
    main()
