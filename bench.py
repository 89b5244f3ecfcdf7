#!/usr/bin/env python3
"""bench.py -- all-vs-all global alignment throughput (GCUPS, pairs/s) on N B200s.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line on rank 0.  A step is one full pass of the hot path (scores + identical counts for
every pair of the upper triangle) over one synthetic protein set.

  value   GCUPS with the packed sequence store already resident in HBM and results left in
          HBM; timed with CUDA events inside the library (first launch -> last kernel end, on
          the streams the kernels run on), summed over the K steps, max over ranks.
  e2e     the same metric through the reference-facing call with HOST buffers: every step
          uploads the raw residues from pinned host memory (bsa_load_sequences), aligns, and
          copies scores + identical counts back to pinned host memory; wall clock.
  roofline  integer-ALU/DPX bound: achieved = cells/s x 7 lane-instructions
          per cell (the kernel's TAG cell) against the lane-op rate of the same
          instruction mix measured live on this GPU by bsa_measure_int_peak.
  cpu_baseline  the oracle (literal C port of the reference path) on the host cores, on a
          bounded random sample of the same workload's pairs.

`--impl reference` times that CPU path alone (the reference itself is Rust and cannot be
built in this image: no rustc/cargo).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_CELL = 7          # lane-instructions of the TAG cell: 4 ALU-pipe + 3 IMAD (gotoh_kernels.cuh)
MATRIX, GO, GE = "BLOSUM62", -10, -1


def workload(n_gpus, override_n=None):
    """configs[1] of BASELINE.json at N=1 (10,000 UniRef50-like proteins).  For N>1 the set grows
    by sqrt(N) so the number of pairs -- the per-GPU work -- stays fixed (weak scaling)."""
    from bioshell_b200 import synth
    n = override_n or int(round(10000 * np.sqrt(n_gpus)))
    cfg = dict(synth.CONFIGS["cfg2"])
    cfg["n"] = n
    res, off = synth.generate(**cfg)
    name = "cfg2: %d synthetic proteins, UniRef50-like lengths (lognormal mu=5.45 sigma=0.65, 30..4000), " \
           "25%% homologs, all-vs-all upper triangle, scores+identity, BLOSUM62 gap -10/-1" % n
    return res, off, name


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            # "under load": drop idle samples taken before the first launch / after the last
            busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": float(max(power))}
        return out


def sample_pairs(n, count, seed):
    rng = np.random.default_rng(seed)
    t = rng.integers(1, n, count)
    q = (rng.random(count) * t).astype(np.int64)
    return q.astype(np.uint32), t.astype(np.uint32)


def cpu_reference_run(res, off, n_pairs, threads, clear_mode, seed=99):
    """The oracle on a bounded sample of the workload's pairs.  Returns (gcups, pairs/s, seconds)."""
    from bioshell_b200.scoring import ncbi_text
    from oracle import c_oracle
    sc, ai = c_oracle.parse_ncbi(ncbi_text(MATRIX))
    S = c_oracle.SeqSet.from_packed(res, off)
    lmax = int(np.diff(off.astype(np.int64)).max())
    q, t = sample_pairs(len(off) - 1, n_pairs, seed)
    t0 = time.perf_counter()
    r = c_oracle.align_pair_list(S, S, sc, ai, GO, GE, q, t, lmax, n_threads=threads, clear_mode=clear_mode)
    dt = time.perf_counter() - t0
    return r["cells"] / 1e9 / dt, n_pairs / dt, dt


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path (literal port) on the host cores."""
    if rank != 0:
        return
    res, off, name = workload(args.gpus, args.n)
    threads = os.cpu_count() or 1
    # size the sample so one step is a few seconds of CPU work
    g0, _, dt0 = cpu_reference_run(res, off, 64 * threads, threads, 0)
    per_step = int(max(64 * threads, min(200000, 64 * threads * 4.0 / max(dt0, 1e-3))))
    for _ in range(args.warmup):
        cpu_reference_run(res, off, max(per_step // 8, threads), threads, 0)
    vals, secs = [], 0.0
    pps = []
    for s in range(args.steps):
        g, p, dt = cpu_reference_run(res, off, per_step, threads, 0, seed=100 + s)
        vals.append(g); pps.append(p); secs += dt
    g_dp, _, _ = cpu_reference_run(res, off, per_step, threads, 1)
    value = float(np.mean(vals))
    sample = "%d random pairs of the workload per step, all %d host threads, faithful per-pair " \
             "(Lmax+1)^2 trace clears (global.rs:69-70)" % (per_step, threads)
    line = {"impl": "reference", "metric": "all-vs-all global alignment throughput", "value": value,
            "unit": "GCUPS", "pairs_per_s": float(np.mean(pps)), "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": name, "timing": "host wall clock, bounded sample extrapolated by cells"},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": "port",
                             "sample": sample, "value_dp_only": g_dp},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="override the number of sequences (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="allvsall", choices=["allvsall", "onevsmany"],
                    help="allvsall = BASELINE configs[1] (the headline, default); onevsmany = configs[3] shape "
                         "(1,000 queries x 125,000 database sequences per GPU, score only, 16-bit lanes)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from bioshell_b200 import Context, SubstitutionMatrix

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the alignment has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner (and any debug output) to
        # stdout unless told otherwise
        if os.environ.get("NCCL_DEBUG"):
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ovm = args.workload == "onevsmany"
    ctx = Context(local_rank)
    ctx.set_scoring(SubstitutionMatrix.load(MATRIX), GO, GE)
    if ovm:
        from bioshell_b200 import synth
        qres, qoff = synth.config("cfg4q")
        res, off = synth.config("cfg4db", n=args.n or 125000 * world)
        n = len(off) - 1
        name = "cfg4 shape: %d queries x %d database sequences (UniRef50-like lengths), score only, " \
               "BLOSUM62 gap -10/-1" % (len(qoff) - 1, n)
        counts = None
        ctx.load_sequences(1, qres, qoff)
        ctx.load_sequences(0, res, off)
        q_set, ops_per_cell, peak_which, want_i = 1, 2.0, 6, False   # 4 ALU-pipe instr per TWO cells (+1 IMAD)
        bounds = ctx.plan_shards(1, 0, None, world)
        t0, t1 = int(bounds[rank]), int(bounds[rank + 1])
        n_res = (t1 - t0) * (len(qoff) - 1)
    else:
        res, off, name = workload(world, args.n)
        n = len(off) - 1
        counts = np.arange(n, dtype=np.uint32)
        ctx.load_sequences(0, res, off)
        q_set, ops_per_cell, peak_which, want_i = 0, OPS_PER_CELL, 7, True
        bounds = ctx.plan_shards(0, 0, counts, world)          # identical on every rank: no exchange
        t0, t1 = int(bounds[rank]), int(bounds[rank + 1])
        n_res = int(counts[t0:t1].astype(np.int64).sum())

    # outputs in HBM for the device-timed leg, pinned host buffers for the e2e leg
    d_scores = torch.empty(max(n_res, 1), dtype=torch.int32, device="cuda")
    d_nid = torch.empty(max(n_res, 1), dtype=torch.int32, device="cuda")
    h_scores = torch.empty(max(n_res, 1), dtype=torch.int32).pin_memory()
    h_nid = torch.empty(max(n_res, 1), dtype=torch.int32).pin_memory()
    h_res = torch.from_numpy(res.copy()).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_device():
        ctx.align_all_pairs(q_set, 0, counts, t0, t1, scores=d_scores.data_ptr(), want_identical=want_i,
                            n_identical=d_nid.data_ptr() if want_i else None, device_out=True)
        return ctx.stats()

    def step_e2e():
        ctx.load_sequences(0, h_res.numpy(), off)
        if ovm:
            ctx.load_sequences(1, qres, qoff)
        ctx.align_all_pairs(q_set, 0, counts, t0, t1, scores=h_scores.numpy(), want_identical=want_i,
                            n_identical=h_nid.numpy() if want_i else None)
        return ctx.stats()

    for _ in range(max(args.warmup, 3)):
        step_device()
    # ---- device-timed leg ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms, launches, cells, pairs, padded = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        flush.zero_()                      # evict L2 between timed iterations
        torch.cuda.synchronize()
        st = step_device()
        dev_ms += st["kernel_ms"]
        launches += st["launches"]
        cells, pairs, padded = st["cells"], st["pairs"], st["padded_cells"]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()

    # ---- end-to-end leg (host buffers in, host buffers out) ----
    step_e2e()
    barrier()
    e0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        st = step_e2e()
        h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3

    peak_ops, peak_mhz = ctx.measure_int_peak(peak_which)

    # whole-job aggregates: max time over ranks, sum of units over ranks
    tv = torch.tensor([dev_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
    uv = torch.tensor([float(cells), float(pairs), float(launches), float(padded)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(uv, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms = tv.tolist()
    cells, pairs, launches, padded = uv.tolist()

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        gcups = cells / 1e9 / (ms_per_step / 1e3)
        e2e_gcups = cells / 1e9 / (e2e_ms / args.steps / 1e3)
        per_gpu_cells_s = gcups * 1e9 / world
        achieved = per_gpu_cells_s * ops_per_cell
        line = {
            "metric": ("one-vs-many" if ovm else "all-vs-all") + " global alignment throughput", "value": gcups, "unit": "GCUPS",
            "pairs_per_s": pairs / (ms_per_step / 1e3), "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16x2" if ovm else "int32", "data": "synthetic",
            "config": {"workload": name, "n_sequences": n, "pairs": int(pairs), "cells": int(cells),
                       "parallelism": "template-range shards x%d, no collective" % world,
                       "l2": "256 MiB buffer rewritten between timed steps (inputs are 2.9 MB; outputs 8 B/pair stream to HBM)",
                       "timing": "CUDA events in the library on the kernels' own streams, summed over steps, max over ranks",
                       "wall_ms_per_step": wall_ms / args.steps,
                       "swept_cells_over_cells": padded / cells if cells else None},
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "bsa_load_sequences from pinned host + bsa_align_all_pairs into pinned host buffers, wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "alu", "achieved": achieved / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
                         "frac": achieved / peak_ops,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the hot kernel
                         # family (gotoh_pair_kernel<20, TAG>, 36.96 ms, 1.16e11 cells) from the ncu --set full
                         # capture in profiles/r1_ncu_summary_tag.md; algorithmic HBM bytes of that
                         # launch are ~8 B/pair of results + the 2.9 MB sequence store
                         "traffic": 3426816,
                         "note": "integer/DPX issue roofline per GPU: cells/s x %.1f lane-instructions per cell vs the same "
                                 "instruction mix (all-vs-all: the TAG cell, VIMNMX3 + LOP3 + 2 VIADDMNMX + 3 IMAD; one-vs-many: "
                                 "ALU-pipe instructions only against the VIADDMNMX.S16x2 rate) in independent chains, measured "
                                 "live by bsa_measure_int_peak (of measured; SM clock %.0f MHz during that probe). HBM is "
                                 "not the bound: algorithmic traffic is 8 B/pair." % (ops_per_cell, peak_mhz),
                         "hbm_algorithmic_gbs": (pairs * 8 / world) / (ms_per_step / 1e3) / 1e9},
        }
        if not args.no_cpu_baseline and world == 1 and not ovm:
            threads = os.cpu_count() or 1
            g0, _, dt0 = cpu_reference_run(res, off, 32 * threads, threads, 0)
            cnt = int(max(32 * threads, min(100000, 32 * threads * 12.0 / max(dt0, 1e-3))))
            g, p, dt = cpu_reference_run(res, off, cnt, threads, 0)
            g_dp, _, _ = cpu_reference_run(res, off, cnt, threads, 1)
            g1, _, _ = cpu_reference_run(res, off, max(cnt // threads, 16), 1, 0)
            line["cpu_baseline"] = {
                "value": g, "unit": "GCUPS", "cores": threads, "kind": "port",
                "sample": "%d random pairs of the same workload (%.1f s), all host threads, faithful to the reference "
                          "incl. its per-pair (Lmax+1)^2 trace clears; value_dp_only clears only the pair's extent; "
                          "value_single_thread is the reference's actual (single-threaded) loop" % (cnt, dt),
                "pairs_per_s": p, "value_dp_only": g_dp, "value_single_thread": g1}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
