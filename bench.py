#!/usr/bin/env python
"""bench.py -- throughput of the BPE-encode hot path (BASELINE.json metric: input GB/s and
Mtokens/s, cl100k_base, 1 GiB synthetic English-like corpus = SURVEY.md 8(d) config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload config2]

A "step" = one pass of the hot path over one batch (the whole workload of this rank).
  value    device-resident: text + doc offsets already in HBM, tokens + offsets left in HBM;
           timed with CUDA events on the launching stream, max over ranks.
  e2e      the same metric through the public host API (Encoding.encode_ordinary_packed ->
           C ABI b200bpe_encode_ordinary_batch) with pinned HOST buffers: H2D of the text, the
           kernels and D2H of tokens + offsets are all inside the timed region.
  roofline achieved algorithmic bytes/s of the dominant kernel (encode_tiles) from CUDA events
           recorded by the engine around that launch, against the measured HBM peak.
  cpu_baseline / --impl reference: the reference engine itself (the tiktoken wheel's Rust CoreBPE
           driven through tiktoken.Encoding.encode_ordinary_batch with all host cores) on a bounded
           sample of the same workload; if the wheel cannot be imported, the oracle port.
N > 1 (torchrun, one rank per GPU): documents shard across ranks (weak scaling: every rank has its
own corpus of the configured size); the only exchange is an NCCL all-gather of per-rank counts.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools import corpus  # noqa: E402
import vocab_util as vu   # noqa: E402

WORKLOADS = {
    # name: (encoding, builder, description)
    "config2": ("cl100k_base", lambda n, seed: corpus.config2(n, seed), "cl100k_base, english-like, ~64 KiB docs"),
    "config3": ("o200k_base", lambda n, seed: corpus.config3(n, seed), "o200k_base, mixed UTF-8, docs 4-256 KiB"),
    "config4": ("cl100k_base", lambda n, seed: corpus.config4(max(1, n // 100), seed), "cl100k_base, ~100 B docs"),
    "config5": ("p50k_base", lambda n, seed: corpus.config5(n, seed), "p50k_base, one code-like document"),
}
DEFAULT_BYTES = {"config2": 1 << 30, "config3": 1 << 30, "config4": 1 << 30, "config5": 64 << 20}
SEEDS = {"config2": 1002, "config3": 1003, "config4": 1004, "config5": 1005}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_reference_engine(pat, ranks, special):
    """The reference engine: installed tiktoken wheel (Rust CoreBPE).  None if not importable."""
    try:
        import tiktoken
        from tiktoken import _tiktoken  # noqa: F401  (make sure it is the native engine)
        return tiktoken.Encoding("bench_ref", pat_str=pat, mergeable_ranks=ranks, special_tokens=special), \
            f"tiktoken=={tiktoken.__version__} wheel (Rust CoreBPE) via Encoding.encode_ordinary_batch"
    except Exception:
        return None, None


def cpu_reference_run(pat, ranks, special, text, off, target_s, cores):
    """Time the reference CPU path on a bounded prefix of the workload (about target_s seconds).
    Returns (GB/s, Mtok/s, description dict)."""
    ref, how = load_reference_engine(pat, ranks, special)
    n_docs = len(off) - 1

    def docs_upto(nbytes):
        k = int(np.searchsorted(off, nbytes, side="right")) - 1
        k = max(1, min(k, n_docs))
        return k, int(off[k])

    if ref is not None:
        kind = "reference"

        def run(k):
            docs = [text[int(off[i]):int(off[i + 1])].tobytes().decode("utf-8") for i in range(k)]
            ref.encode("warmup")
            t0 = time.perf_counter()
            out = ref.encode_ordinary_batch(docs, num_threads=cores)
            dt = time.perf_counter() - t0
            return dt, sum(len(x) for x in out)
    else:
        from oracle import Oracle
        orc = Oracle(ranks, special, pat)
        kind, how = "port", "oracle/bpe_oracle.c (C restatement), pthread batch driver"

        def run(k):
            t0 = time.perf_counter()
            toks, _ = orc.encode_ordinary_batch_np(text[:int(off[k])], off[:k + 1], cores)
            return time.perf_counter() - t0, len(toks)

    k0, b0 = docs_upto(8 << 20)
    dt0, _ = run(k0)                                          # probe to size the sample
    rate = b0 / max(dt0, 1e-6)
    k, b = docs_upto(int(min(len(text), max(b0, rate * target_s))))
    dt, ntok = run(k)
    return b / dt / 1e9, ntok / dt / 1e6, {"kind": kind, "cores": cores, "how": how, "seconds": dt,
                                           "sample": f"first {k} docs = {b} bytes of the workload, {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--bytes", type=int, default=0, help="override the per-rank corpus size (development)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode", action="store_true", help="also time the device decode of the produced tokens (next row)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    enc_name, builder, wl_desc = WORKLOADS[args.workload]
    nbytes = args.bytes or DEFAULT_BYTES[args.workload]
    pat, ranks, special, vocab_src = vu.load_encoding(enc_name)
    cores = os.cpu_count() or 1
    config = {"workload": f"{args.workload}: {wl_desc}", "bytes_per_gpu": nbytes, "encoding": enc_name,
              "vocab": f"{vocab_src} ({len(ranks)} mergeable ranks)", "seed": SEEDS[args.workload],
              "l2": "inputs (>= 64 MiB text per step, streamed once) exceed or equal the 126 MB L2; no reuse between steps",
              "parallelism": f"doc-sharded x{world}"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        text, off = builder(min(nbytes, 256 << 20), SEEDS[args.workload])
        vals, toks, secs = [], [], []
        desc = None
        for i in range(args.warmup + args.steps):
            gbs, mts, desc = cpu_reference_run(pat, ranks, special, text, off, args.cpu_seconds, cores)
            if i >= args.warmup:
                vals.append(gbs); toks.append(mts); secs.append(desc["seconds"])
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "input_GB_per_s", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "mtokens_per_s": float(np.mean(toks)), "config": config,
                "cpu_baseline": {"value": v, "unit": "GB/s", **desc},
                "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: tiktoken_b200 has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import tiktoken_b200
    from tiktoken_b200.sharding import CountExchange, gather_counts

    enc = tiktoken_b200.Encoding(enc_name + "_bench", pat_str=pat, mergeable_ranks=ranks, special_tokens=special,
                                 device=local_rank)
    core = enc._core_bpe
    text, off = builder(nbytes, SEEDS[args.workload] + 7919 * rank)        # weak scaling: own corpus per rank
    n_docs = len(off) - 1
    N = len(text)

    # pinned host copies (the e2e path copies FROM these every step)
    h_text = torch.empty(N, dtype=torch.uint8, pin_memory=True)
    h_text.numpy()[:] = text
    h_off = torch.empty(n_docs + 1, dtype=torch.int64, pin_memory=True)
    h_off.numpy()[:] = off.astype(np.int64)
    # device-resident inputs / outputs for `value`
    d_text = h_text.cuda(non_blocking=True)
    d_off = h_off.cuda(non_blocking=True)
    d_tok = torch.empty(N, dtype=torch.int32, device="cuda")
    d_toff = torch.empty(n_docs + 1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def step_device():
        return core.encode_device(d_text.data_ptr(), N, d_off.data_ptr(), n_docs, d_tok.data_ptr(), d_toff.data_ptr(),
                                  stream.cuda_stream)

    # ---- parity gate on a sample before any number is reported
    n_tok = step_device()
    if rank == 0:
        from oracle import Oracle
        orc = Oracle(ranks, special, pat)
        k = int(min(n_docs, max(1, np.searchsorted(off, 6 << 20))))
        exp_t, exp_o = orc.encode_ordinary_batch_np(text[:int(off[k])], off[:k + 1], cores)
        got_o = d_toff[:k + 1].cpu().numpy().astype(np.uint64)
        got_t = d_tok[:int(got_o[-1])].cpu().numpy().view(np.uint32)
        if not (np.array_equal(got_o, exp_o) and np.array_equal(got_t, exp_t)):
            print(json.dumps({"error": "PARITY FAILURE against the oracle on the bench workload"}))
            return 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(args.warmup):
        gather_counts(step_device(), n_docs, rank, world, device="cuda")     # also warms the NCCL communicator
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"pretok_ms": [], "encode_ms": [], "probe_ms": [], "gather_ms": [], "long_ms": [], "mark_docs_ms": [], "device_total_ms": []}
    launches = 0
    ev0.record(stream)
    xchg = CountExchange(rank, world, device="cuda")
    for _ in range(args.steps):
        n_tok = step_device()
        xchg.post(n_tok, n_docs)           # NCCL all-gather of (tokens, docs), next to the following step's kernels
        tm = core.last_timings()
        for key in stage:
            stage[key].append(tm[key])
        launches += tm["launches"]
    placements = xchg.drain()              # every exchange completes inside the timed region
    counts = placements[-1][0]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    tot = torch.tensor([N, n_tok], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tot_bytes, tot_tokens = int(tot[0].item()), int(tot[1].item())
    value = tot_bytes / (ms_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the public API, copies inside the timed region
    h_text_np, h_off_np = h_text.numpy(), h_off.numpy().view(np.uint64)
    for _ in range(2):
        enc.encode_ordinary_packed(h_text_np, h_off_np).close()
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        buf = enc.encode_ordinary_packed(h_text_np, h_off_np)      # H2D + kernels + D2H, synchronous
        e2e_t.append(time.perf_counter() - t0)
        e2e_tokens = buf.n_tokens
        buf.close()
    barrier()
    t2 = torch.tensor([float(np.mean(e2e_t))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = tot_bytes / float(t2.item()) / 1e9
    e2e_tm = core.last_timings()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (probe_kernel), algorithmic bytes per launch
    peak, peak_src = measured_peak()
    probe_ms = float(np.mean(stage["probe_ms"]))
    enc_ms = float(np.mean(stage["encode_ms"]))
    alg_probe = N + N // 8 + 4 * n_tok                      # text + piece bitmask read, one 4-byte slot per piece written
    achieved = alg_probe / (probe_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("probe_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    pre_ms = float(np.mean(stage["pretok_ms"]))
    pipeline_alg = N + 4 * n_tok + 16 * (n_docs + 1)
    dev_ms = float(np.mean(stage["device_total_ms"]))
    line = {
        "metric": "input_GB_per_s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
        "mtokens_per_s": tot_tokens / (ms_step * 1e-3) / 1e6, "bytes_per_token": tot_bytes / max(1, tot_tokens),
        "n_docs_per_gpu": n_docs, "gpu_launches": launches,
        "stage_ms": {k: float(np.mean(v)) for k, v in stage.items()},
        "roofline": {"bound": "hbm", "kernel": "probe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_probe, "kernel_ms": probe_ms,
                     "encode_stage": {"kernels": "probe + miss sort + miss merge", "ms": enc_ms,
                                      "achieved": (N + N // 8 + 4 * n_tok) / (enc_ms * 1e-3) / 1e9},
                     "pretok_kernel": {"achieved": (N + N // 8 + N // 8) / (pre_ms * 1e-3) / 1e9, "kernel_ms": pre_ms,
                                       "frac": (N + N // 4) / (pre_ms * 1e-3) / 1e9 / peak},
                     "pipeline": {"achieved": pipeline_alg / (dev_ms * 1e-3) / 1e9,
                                  "frac": pipeline_alg / (dev_ms * 1e-3) / 1e9 / peak,
                                  "hbm_read_only_frac": N / (dev_ms * 1e-3) / 1e9 / peak, "device_ms": dev_ms}},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(N + 8 * (n_docs + 1)),
                "d2h_bytes_per_step": int(4 * e2e_tokens + 8 * (n_docs + 1)), "ms_per_step": float(t2.item()) * 1e3,
                "mtokens_per_s": tot_tokens / float(t2.item()) / 1e6,
                "h2d_ms": e2e_tm["h2d_ms"], "d2h_ms": e2e_tm["d2h_ms"], "device_ms": e2e_tm["device_total_ms"]},
        "clocks": clocks,
    }
    if args.decode and world == 1:
        buf = enc.encode_ordinary_packed(h_text_np, h_off_np)
        dtoks, doffs = np.array(buf.tokens()), np.array(buf.offsets())
        buf.close()
        enc.decode_packed(dtoks, doffs)
        dt = []
        for _ in range(3):
            t0 = time.perf_counter()
            data, boff = enc.decode_packed(dtoks, doffs)
            dt.append(time.perf_counter() - t0)
        assert len(data) == N
        line["decode"] = {"value": N / float(np.mean(dt)) / 1e9, "unit": "GB/s of decoded bytes (host tokens -> host bytes)",
                          "device_ms": core.last_timings()["device_total_ms"], "ms_per_step": float(np.mean(dt)) * 1e3}
    if not args.no_cpu_baseline and world == 1:
        gbs, mts, desc = cpu_reference_run(pat, ranks, special, text, off, args.cpu_seconds, cores)
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "mtokens_per_s": mts, **desc}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
