#!/usr/bin/env python3
"""bench.py -- global-alignment throughput (GCUPS, pairs/s) on N B200s, one line per workload.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line on rank 0.  A step is one full pass of the hot path over one synthetic input.

`--workload` names the BASELINE.json configuration; the default is the one the
metric is quoted on:
  cfg2       configs[1]  10,000 UniRef50-like proteins, all-vs-all, scores + identity   (DEFAULT)
  cfg1       configs[0]  1,000 proteins (len 50-500), all-vs-all, scores + identity
  cfg3shard  configs[2]  100,000 proteins: rank r computes template-range shard r of max(8, N)
                         (N = 8 is the whole job, N = 1 is one eighth of it)
  cfg4       configs[3]  1,000 queries x 125,000 database sequences PER GPU, score only
                         (N = 8 is the whole 1,000 x 1,000,000 job)
  cfg5       configs[4]  16 titin-scale pairs (5k-35k residues, one 34,350 x 35,000) with full
                         traceback through the wavefront kernel

  value   GCUPS with the packed sequence store already resident in HBM and results left in
          HBM; timed with CUDA events inside the library (first launch -> last kernel end, on
          the streams the kernels run on), summed over the K steps, max over ranks.
  e2e     the same metric through the reference-facing call with HOST buffers: every step
          uploads the raw residues from pinned host memory (bsa_load_sequences), aligns, and
          copies the results back to pinned host memory; wall clock.
  roofline  integer-ALU/DPX issue bound: achieved = cells/s x lane-instructions
          per cell of the workload's dominant kernel against the lane-op rate of the same
          instruction mix measured live on this GPU by bsa_measure_int_peak.
  cpu_baseline  the oracle (literal C port of the reference path, rebuilt -O3 -march=native on
          this host) on the host cores, on a bounded random sample of the same workload's pairs.

`--single-process` (with --gpus N, NOT under torchrun): ONE process drives all N GPUs through
the library's own multi-device context (bsa_create_multi: a host worker thread per GPU, results
streamed tile-wise into the caller's pinned buffer) -- the mode a single-process caller such as
bin/cluster_sequences.rs would use.

`--impl reference` times the CPU path alone (the reference itself is Rust and cannot be
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

MATRIX, GO, GE = "BLOSUM62", -10, -1
LAW = "UniRef50-like lengths (lognormal mu=5.45 sigma=0.65, 30..4000), 25% homologs"

# lane-instructions per cell of the dominant kernel and the bsa_measure_int_peak mix they are held against
ROOF = {
    # frame cell with column-tagged E openings (TAG kernels, round 2): 4 ALU-pipe + 1 IMAD (gotoh_stream_kernel /
    # gotoh_pair_kernel; the multi-pass groups, 3 % of cfg2's cells, still run the 6-instruction form)
    "tag": dict(ops=5.0, which=10, mix="frame cell, E openings tagged by column: VIMNMX3 + LOP3 + 2 VIADDMNMX + IMAD",
                alu_ops=4.0),
    # 16-bit packed score-only cell in the moving frame: 4 DPX instructions per TWO cells, nothing else (gotoh_score16_kernel)
    "s16": dict(ops=2.0, which=6, mix="u16x2 frame cell (3 VIADDMNMX.U16x2 + VIMNMX.U16x2 per two cells) against the VIADDMNMX.S16x2 rate",
                alu_ops=2.0),
    # K3 direction-frame cell (gotoh_wave_kernel, round 2): 8 ALU-pipe (VIMNMX3, 4 LOP3, 2 VIADDMNMX, SHF) + 3 IMAD
    "dirs": dict(ops=11.0, which=9, mix="K3 direction-frame cell: VIMNMX3 + 4 LOP3 + 2 VIADDMNMX + SHF + 3 IMAD",
                 alu_ops=8.0),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu --set full
# captures committed under profiles/ (quoted, not re-measured per run: counters need a profiler)
TRAFFIC_QUOTED = {
    "tag": dict(bytes=3194112, source="profiles/r2_ncu_summary_k1_aligned.md: gotoh_stream_kernel<17,single,TAG>, 41.8 ms launch of the cfg2 run"),
    "s16": dict(bytes=1360128, source="profiles/r2_ncu_summary_k1s_frame.md: gotoh_score16_quad_kernel<15>, 22.1 ms launch of a 1000 x 50000 run"),
    "dirs": dict(bytes=4381725952, source="profiles/r2_ncu_summary_wave_frame.md: gotoh_wave_kernel on cfg5, 4.20 GB of directions written + 0.18 GB read"),
}


class Workload:
    """What one rank aligns in a step, and the pairs a CPU sample is drawn from."""
    kind = "allpairs"          # "allpairs" (bsa_align_all_pairs) or "paths" (bsa_align_pairs_paths)
    roof = "tag"
    want_i = True
    scaling = "weak"
    qres = qoff = None         # separate query set (one-vs-many)

    def sample(self, count, seed):
        """-> (q idx, t idx) of `count` random pairs of this workload."""
        n = len(self.off) - 1
        rng = np.random.default_rng(seed)
        if self.qoff is not None:
            return (rng.integers(0, len(self.qoff) - 1, count).astype(np.uint32),
                    rng.integers(0, n, count).astype(np.uint32))
        t = rng.integers(1, n, count)
        q = (rng.random(count) * t).astype(np.int64)
        return q.astype(np.uint32), t.astype(np.uint32)


def make_workload(name, world, rank, override_n=None):
    from bioshell_b200 import synth
    w = Workload()
    w.key = name
    if name in ("cfg1", "cfg2", "cfg3shard"):
        base = {"cfg1": "cfg1", "cfg2": "cfg2", "cfg3shard": "cfg3"}[name]
        cfg = dict(synth.CONFIGS[base])
        if override_n:
            cfg["n"] = override_n
        elif name == "cfg2":
            # weak scaling: the set grows by sqrt(N) so that the pairs per GPU stay fixed
            cfg["n"] = int(round(10000 * np.sqrt(world)))
        w.res, w.off = synth.generate(**cfg)
        n = cfg["n"]
        w.counts = np.arange(n, dtype=np.uint32)
        w.n_shards = max(8, world) if name == "cfg3shard" else world
        if name == "cfg1":
            w.name = "cfg1: %d synthetic proteins, lengths U{50..500}, 25%% homologs, all-vs-all upper triangle, " \
                     "scores+identity, BLOSUM62 gap -10/-1" % n
        elif name == "cfg2":
            w.name = "cfg2: %d synthetic proteins, %s, all-vs-all upper triangle, scores+identity, " \
                     "BLOSUM62 gap -10/-1" % (n, LAW)
        else:
            w.name = "cfg3 shard: %d synthetic proteins, %s, all-vs-all upper triangle, scores+identity, BLOSUM62 gap " \
                     "-10/-1; each GPU computes one cell-balanced template-range shard of %d (%d of %d shards run)" % (
                         n, LAW, w.n_shards, world, w.n_shards)
    elif name == "cfg4":
        w.qres, w.qoff = synth.config("cfg4q")
        w.res, w.off = synth.config("cfg4db", n=override_n or 125000 * world)
        w.counts = None
        w.n_shards = world
        w.roof, w.want_i = "s16", False
        w.name = "cfg4 shape: %d queries x %d database sequences (%s), score only, BLOSUM62 gap -10/-1" % (
            len(w.qoff) - 1, len(w.off) - 1, LAW)
    elif name == "cfg5":
        cfg = dict(synth.CONFIGS["cfg5"])
        cfg["seed"] += rank          # N > 1: every GPU aligns its own 16 pairs
        if override_n:
            cfg["pairs"] = override_n
        w.res, w.off = synth.pair_set(**cfg)
        w.kind, w.roof = "paths", "dirs"
        w.pq = np.arange(0, 2 * cfg["pairs"], 2, dtype=np.uint32)
        w.pt = w.pq + 1
        lens = np.diff(w.off.astype(np.int64))
        w.name = "cfg5: %d titin-scale pairs per GPU (lengths %d..%d, even pairs homologous, pair 0 = %d x %d), " \
                 "full traceback strings, BLOSUM62 gap -10/-1" % (cfg["pairs"], lens.min(), lens.max(), lens[0], lens[1])
        w.sample = lambda count, seed: (w.pq[2:2 + count], w.pt[2:2 + count])     # pairs 2.. (13k-16k residues)
    else:
        raise SystemExit("unknown workload " + name)
    return w


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


def cpu_reference_run(w, n_pairs, threads, clear_mode, seed=99):
    """The oracle on a bounded sample of the workload's pairs.  Returns (gcups, pairs/s, seconds, pairs)."""
    from bioshell_b200.scoring import ncbi_text
    from oracle import c_oracle
    c_oracle.use_native()            # -O3 -march=native build for THIS host
    sc, ai = c_oracle.parse_ncbi(ncbi_text(MATRIX))
    T = c_oracle.SeqSet.from_packed(w.res, w.off)
    Q = T if w.qoff is None else c_oracle.SeqSet.from_packed(w.qres, w.qoff)
    lmax = int(np.diff(w.off.astype(np.int64)).max())
    if w.qoff is not None:
        lmax = max(lmax, int(np.diff(w.qoff.astype(np.int64)).max()))
    q, t = w.sample(n_pairs, seed)
    if w.kind == "paths":
        # the aligner is sized for the sampled pairs only (3 x (Lmax+1)^2 bytes per thread: 3.7 GB at 35,000)
        ln = np.diff(w.off.astype(np.int64))
        lmax = int(max(ln[q.astype(np.int64)].max(), ln[t.astype(np.int64)].max()))
    t0 = time.perf_counter()
    # score-only workloads still run the reference's whole loop body (align + backtrace + report)
    r = c_oracle.align_pair_list(Q, T, sc, ai, GO, GE, q, t, lmax, n_threads=threads, clear_mode=clear_mode)
    dt = time.perf_counter() - t0
    return r["cells"] / 1e9 / dt, len(q) / dt, dt, len(q)


def cpu_baseline(w, threads, budget_s):
    """Bounded CPU sample: faithful (per-pair (Lmax+1)^2 trace clears, global.rs:69-70), DP-only, single thread."""
    if w.kind == "paths":
        # whole titin-scale pairs: one oracle call per pair, pairs in parallel threads
        cnt = min(2, len(w.pq) - 2)
        g, p, dt, cnt = cpu_reference_run(w, cnt, min(threads, cnt), 0)
        g1, _, _, _ = cpu_reference_run(w, 1, 1, 0)
        return dict(value=g, pairs_per_s=p, value_dp_only=g, value_single_thread=g1, pairs=cnt, seconds=dt,
                    cores=min(threads, cnt))
    _, _, dt0, _ = cpu_reference_run(w, 32 * threads, threads, 0)
    cnt = int(max(32 * threads, min(100000, 32 * threads * budget_s / max(dt0, 1e-3))))
    g, p, dt, _ = cpu_reference_run(w, cnt, threads, 0)
    g_dp, _, _, _ = cpu_reference_run(w, cnt, threads, 1)
    g1, _, _, _ = cpu_reference_run(w, max(cnt // threads, 16), 1, 0)
    return dict(value=g, pairs_per_s=p, value_dp_only=g_dp, value_single_thread=g1, pairs=cnt, seconds=dt, cores=threads)


def cpu_baseline_json(b):
    return {"value": b["value"], "unit": "GCUPS", "cores": b["cores"], "kind": "port",
            "sample": "%d random pairs of the same workload (%.1f s), %d host threads, oracle rebuilt -O3 -march=native on "
                      "this host, faithful to the reference incl. its per-pair (Lmax+1)^2 trace clears; value_dp_only clears "
                      "only the pair's extent; value_single_thread is the reference's actual (single-threaded) loop"
                      % (b["pairs"], b["seconds"], b["cores"]),
            "pairs_per_s": b["pairs_per_s"], "value_dp_only": b["value_dp_only"],
            "value_single_thread": b["value_single_thread"]}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path (literal port) on the host cores."""
    if rank != 0:
        return
    w = make_workload(args.workload, args.gpus, 0, args.n)
    threads = os.cpu_count() or 1
    if w.kind == "paths":
        per_step, thr = min(2, len(w.pq) - 2), min(threads, 2)
    else:
        _, _, dt0, _ = cpu_reference_run(w, 64 * threads, threads, 0)
        per_step, thr = int(max(64 * threads, min(200000, 64 * threads * 4.0 / max(dt0, 1e-3)))), threads
    for _ in range(args.warmup):
        cpu_reference_run(w, max(per_step // 8, 1), thr, 0)
    vals, secs, pps = [], 0.0, []
    for s in range(args.steps):
        g, p, dt, _ = cpu_reference_run(w, per_step, thr, 0, seed=100 + s)
        vals.append(g); pps.append(p); secs += dt
    g_dp, _, _, _ = cpu_reference_run(w, per_step, thr, 1)
    g1, _, _, _ = cpu_reference_run(w, max(per_step // thr, 1), 1, 0)
    value = float(np.mean(vals))
    sample = "%d pairs of the workload per step, %d host threads, oracle rebuilt -O3 -march=native on this host, faithful " \
             "per-pair (Lmax+1)^2 trace clears (global.rs:69-70)" % (per_step, thr)
    line = {"impl": "reference", "metric": metric_name(w), "value": value,
            "unit": "GCUPS", "pairs_per_s": float(np.mean(pps)), "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": w.scaling, "vs_baseline": None, "dtype": dtype_name(w), "data": "synthetic",
            "config": {"workload": w.name, "timing": "host wall clock, bounded sample extrapolated by cells"},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": thr, "kind": "port",
                             "sample": sample, "value_dp_only": g_dp, "value_single_thread": g1},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def metric_name(w):
    return {"cfg4": "one-vs-many", "cfg5": "long-pair full-traceback"}.get(w.key, "all-vs-all") + \
        " global alignment throughput"


def dtype_name(w):
    return "int16x2" if w.roof == "s16" else "int32"


class Runner:
    """One rank's (or, single-process, all GPUs') step functions over a context."""

    def __init__(self, ctx, w, rank, world, torch, multi=False):
        self.ctx, self.w, self.torch = ctx, w, torch
        ctx.set_scoring(MATRIX, GO, GE)
        self.h_res = torch.from_numpy(w.res.copy()).pin_memory()
        if w.qoff is not None:
            ctx.load_sequences(1, w.qres, w.qoff)
        ctx.load_sequences(0, w.res, w.off)
        self.q_set = 1 if w.qoff is not None else 0
        self.multi = multi
        if w.kind == "allpairs":
            nq = (len(w.qoff) if w.qoff is not None else len(w.off)) - 1
            if multi:
                self.t0, self.t1 = 0, len(w.off) - 1
            else:
                bounds = ctx.plan_shards(self.q_set, 0, w.counts, w.n_shards)      # identical on every rank: no exchange
                self.t0, self.t1 = int(bounds[rank]), int(bounds[rank + 1])
            self.n_res = int(w.counts[self.t0:self.t1].astype(np.int64).sum()) if w.counts is not None \
                else (self.t1 - self.t0) * nq
            n_alloc = max(self.n_res, 1)
            # outputs in HBM for the device-timed leg, pinned host buffers for the e2e leg
            if not multi:
                self.d_scores = torch.empty(n_alloc, dtype=torch.int32, device="cuda")
                self.d_nid = torch.empty(n_alloc, dtype=torch.int32, device="cuda") if w.want_i else None
            self.h_scores = torch.empty(n_alloc, dtype=torch.int32).pin_memory()
            self.h_nid = torch.empty(n_alloc, dtype=torch.int32).pin_memory() if w.want_i else None

    def step_device(self):
        w, ctx = self.w, self.ctx
        if w.kind == "paths":
            ctx.align_pairs_paths(0, 0, w.pq, w.pt)
        elif self.multi:
            # one process, every GPU: results can only land in host memory
            ctx.align_all_pairs(self.q_set, 0, w.counts, self.t0, self.t1, scores=self.h_scores.numpy(),
                                want_identical=w.want_i, n_identical=self.h_nid.numpy() if w.want_i else None)
        else:
            ctx.align_all_pairs(self.q_set, 0, w.counts, self.t0, self.t1, scores=self.d_scores.data_ptr(),
                                want_identical=w.want_i, n_identical=self.d_nid.data_ptr() if w.want_i else None,
                                device_out=True)
        return ctx.stats()

    def step_e2e(self):
        w, ctx = self.w, self.ctx
        ctx.load_sequences(0, self.h_res.numpy(), w.off)
        if w.qoff is not None:
            ctx.load_sequences(1, w.qres, w.qoff)
        if w.kind == "paths":
            ctx.align_pairs_paths(0, 0, w.pq, w.pt)
        else:
            ctx.align_all_pairs(self.q_set, 0, w.counts, self.t0, self.t1, scores=self.h_scores.numpy(),
                                want_identical=w.want_i, n_identical=self.h_nid.numpy() if w.want_i else None)
        return ctx.stats()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="override the number of sequences / pairs (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2",
                    choices=["cfg1", "cfg2", "cfg3shard", "cfg4", "cfg5", "allvsall", "onevsmany"],
                    help="BASELINE.json configuration (see the module docstring); cfg2 is the headline and the default")
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives all --gpus N devices through bsa_create_multi (do not use torchrun)")
    args = ap.parse_args()
    args.workload = {"allvsall": "cfg2", "onevsmany": "cfg4"}.get(args.workload, args.workload)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from bioshell_b200 import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the alignment has no CPU fallback")
    single = args.single_process
    if single and world > 1:
        raise SystemExit("--single-process is one process for all GPUs: run it without torchrun")
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
        if single:
            for d in range(args.gpus):
                torch.cuda.synchronize(d)
        else:
            torch.cuda.synchronize()

    n_dev = args.gpus if single else world
    w = make_workload(args.workload, n_dev, rank, args.n)
    if single:
        if w.kind != "allpairs":
            raise SystemExit("--single-process covers the all-pairs workloads (cfg1..cfg4)")
        w.n_shards = 1
        ctx = Context(list(range(args.gpus)))
    else:
        ctx = Context(local_rank)
    run = Runner(ctx, w, rank, world, torch, multi=single)
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % d)
             for d in (range(args.gpus) if single else [local_rank])]           # > 126 MB L2

    for _ in range(max(args.warmup, 3)):
        run.step_device()
    # ---- device-timed leg ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms, launches, cells, pairs, padded = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        for f in flush:
            f.zero_()                      # evict L2 between timed iterations
        barrier()
        st = run.step_device()
        dev_ms += st["kernel_ms"]
        launches += st["launches"]
        cells, pairs, padded = st["cells"], st["pairs"], st["padded_cells"]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()

    # ---- end-to-end leg (host buffers in, host buffers out) ----
    run.step_e2e()
    barrier()
    e0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        st = run.step_e2e()
        h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3

    roof = ROOF[w.roof]
    peak_ops, peak_mhz = ctx.measure_int_peak(roof["which"])
    alu_rate, _ = ctx.measure_int_peak(1)      # VIADDMNMX alone: the rate of the ALU pipe that carries the DPX instructions

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
        per_gpu_cells_s = gcups * 1e9 / n_dev
        achieved = per_gpu_cells_s * roof["ops"]
        res_bytes = {"allpairs": 8 if w.want_i else 4, "paths": 8}[w.kind]
        tq = TRAFFIC_QUOTED[w.roof]
        n_seq = len(w.off) - 1
        line = {
            "metric": metric_name(w), "value": gcups, "unit": "GCUPS",
            "pairs_per_s": pairs / (ms_per_step / 1e3), "n_gpus": n_dev, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": w.scaling, "vs_baseline": None, "dtype": dtype_name(w), "data": "synthetic",
            "config": {"workload": w.name, "n_sequences": n_seq, "pairs": int(pairs), "cells": int(cells),
                       "parallelism": ("one process, %d GPUs inside the library (bsa_create_multi), no collective" % n_dev) if single
                       else "template-range shards x%d, one process per GPU, no collective" % world,
                       "l2": "256 MiB buffer rewritten between timed steps (inputs are a few MB; results stream to HBM)",
                       "timing": ("wall clock of the multi-device call inside the library (results land in pinned host memory)" if single else
                                  "CUDA events in the library on the kernels' own streams, summed over steps, max over ranks"),
                       "wall_ms_per_step": wall_ms / args.steps,
                       "swept_cells_over_cells": padded / cells if cells and padded else None},
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "bsa_load_sequences from pinned host + the alignment call into (pinned) host buffers, wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "alu", "achieved": achieved / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
                         "frac": achieved / peak_ops,
                         # the second, mix-independent denominator: the ALU-pipe instructions of the cell alone against the
                         # measured rate of that pipe (a bound no instruction mix can beat)
                         "alu_pipe": {"achieved": per_gpu_cells_s * roof["alu_ops"] / 1e12, "peak": alu_rate / 1e12,
                                      "unit": "Tlane-op/s", "frac": per_gpu_cells_s * roof["alu_ops"] / alu_rate,
                                      "ops_per_cell": roof["alu_ops"]},
                         "traffic": tq["bytes"], "traffic_source": tq["source"],
                         "note": "integer/DPX issue roofline per GPU: cells/s x %.1f lane-instructions per cell (%s) in independent "
                                 "chains, measured live by bsa_measure_int_peak (of measured; SM clock %.0f MHz during that probe). "
                                 "HBM is not the bound: algorithmic traffic is %d B/pair%s. `traffic` is quoted from the committed ncu "
                                 "capture named in traffic_source (DRAM counters need a profiler), not re-measured in this run."
                                 % (roof["ops"], roof["mix"], peak_mhz, res_bytes,
                                    " + 0.5 B/cell of directions" if w.kind == "paths" else ""),
                         "hbm_algorithmic_gbs": ((pairs * res_bytes + (cells * 0.5 if w.kind == "paths" else 0)) / n_dev)
                         / (ms_per_step / 1e3) / 1e9},
        }
        if not args.no_cpu_baseline and n_dev == 1:
            line["cpu_baseline"] = cpu_baseline_json(cpu_baseline(w, os.cpu_count() or 1, 12.0))
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
