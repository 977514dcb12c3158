#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on synthetic data.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
                    [--config C1|C2|C3|C4-maxmean|C4-mean|C5]

metric: gene-set x permutation ES/sec of the sample-permutation path (permute ->
gene score -> rank -> ES -> NES / p / FDR), config C2 by default: flexgsea_s2n +
flexgsea_weighted_ks (p=1), 20,000 genes x 200 samples two-class, 8,000 gene sets
of 15-500 genes, nperm=10,000.  One "step" = the null loop over the config's
permutations plus flexgsea_calc_sig of every response.

  value : S*R*P / device time (CUDA events on the library's stream from the
          first kernel of the null loop to the last of the significance, inputs
          resident in HBM, permutations drawn on the device), max over ranks.
  e2e   : the same metric through the public C-ABI calls a flexgsea() call
          makes, host buffers in pinned memory, H2D/D2H inside the timed
          region, wall clock.
  roofline / cpu_baseline / clocks / stages: see DESIGN.md "Measurement".

N > 1 (torchrun): STRONG scaling -- the config's nperm is sharded, rank r takes
permutations [r*P/N, (r+1)*P/N) (Philox keyed by the global permutation id, so
the null does not depend on N); the significance exchange (NCCL, inside the
library: fgs_calc_sig_sharded) is in the timed region and its time is reported
as collective_ms_per_step.
--impl reference: the reference's CPU path.  R is not installable here (no R,
no network), so this arm times the oracle restatement of the reference's loop
(oracle/oracle.c, OpenMP over permutations, all host cores) on a bounded
sample of the same workload; rank 0 only.
"""
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

WORKLOAD_TEXT = {
    "C1": "C1: s2n + weighted KS (p=1), 1,000 genes x 20 samples two-class, 50 gene sets (10-50), nperm=1,000",
    "C2": "C2: s2n + weighted KS (p=1), 20,000 genes x 200 samples two-class, 8,000 gene sets (15-500), nperm=10,000",
    "C3": "C3: flexgsea_lm (response + 3 covariates, 4 score columns) + weighted KS, 20,000 genes x 500 samples, "
          "8,000 gene sets (15-500), nperm=10,000",
    "C4-maxmean": "C4: s2n + maxmean ES with NES-based FDR, 20,000 genes x 200 samples two-class, 20,000 gene sets "
                  "(15-300), nperm=50,000",
    "C4-mean": "C4: s2n + mean ES with NES-based FDR, 20,000 genes x 200 samples two-class, 20,000 gene sets "
               "(15-300), nperm=50,000",
    "C5": "C5: s2n + weighted KS, 60,000 transcripts x 1,000 samples two-class, 20,000 gene sets (15-500), "
          "nperm=100,000",
}
WORKLOAD_TEXT["C4"] = WORKLOAD_TEXT["C4-maxmean"]
# config -> (synthetic shape in synth.CONFIGS, gene score, ES)
KINDS = {"C1": ("C1", "s2n", "wks"), "C2": ("C2", "s2n", "wks"), "C3": ("C3", "lm", "wks"),
         "C4": ("C4", "s2n", "maxmean"), "C4-maxmean": ("C4", "s2n", "maxmean"), "C4-mean": ("C4", "s2n", "mean"),
         "C5": ("C5", "s2n", "wks")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The process is started before the
    warm-up (it needs a few hundred ms to enumerate the box) and every line is stamped on arrival; stop() keeps the
    samples that arrived between begin() and stop().  When the timed region is shorter than one sampling period
    and holds none, the samples of the warm-up (same work, same load) are reported and "window" says so."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0 = index, None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = [ln for (ts, ln) in self.lines if self.t0 is None or self.t0 <= ts <= t1]
        window = "timed" if timed else "warmup"
        for ln in (timed or [ln for (_, ln) in self.lines]):
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def workload(config, seed=1):
    """(shape dict, x, response, has_intercept, offsets, index) of a BASELINE config;
    response = y_bin [n x 1] int32 (s2n) or the model matrix [n x 5] (lm: intercept,
    response, three covariates -> R = 4 score columns)."""
    from flexgsea_r_b200 import synth
    shape, score, _ = KINDS[config]
    c = dict(synth.CONFIGS[shape])
    off, idx = synth.gene_sets_csr(c["G"], c["S"], c["sizes"], c["size_dist"], seed=seed)
    x, y = synth.two_class(c["G"], c["n"], off, idx, seed=seed + 1)
    if score == "lm":
        return c, x, synth.continuous_response(c["n"], k=3, seed=seed + 2), True, off, idx
    return c, x, y.astype(np.int32), False, off, idx


def pinned(a):
    """copy of `a` in page-locked host memory (column-major kept)"""
    import torch
    flat = np.asfortranarray(a).reshape(-1, order="F")
    t = torch.from_numpy(flat.copy()).pin_memory()
    return t.numpy().reshape(a.shape, order="F"), t


def oracle_kinds(orc, config):
    _, score, es = KINDS[config]
    return (orc.SCORE_LM if score == "lm" else orc.SCORE_S2N,
            {"wks": orc.ES_WKS, "maxmean": orc.ES_MAXMEAN, "mean": orc.ES_MEAN}[es])


def cpu_sample_perms(config, cores):
    """permutations per CPU step: about 10-30 s of work on the box's cores"""
    base = {"C1": 1000, "C2": 8, "C3": 2, "C4": 8, "C4-maxmean": 8, "C4-mean": 8, "C5": 1}[config]
    return max(base * cores, 16)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    from flexgsea_r_b200 import synth
    orc.build()
    c, x, resp, icpt, off, idx = workload(args.config)
    sk, ek = oracle_kinds(orc, args.config)
    R = resp.shape[1] - 1 if sk == orc.SCORE_LM else resp.shape[1]
    cores = orc.use_all_cores()
    P_s = args.ref_perms if args.ref_perms else cpu_sample_perms(args.config, cores)
    perm = synth.perm_matrix(c["n"], P_s, seed=4)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc, _ = orc.perm_null(x, sk, resp, icpt, perm, off, idx, ek)
        dt = time.perf_counter() - t0
        assert rc == 0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = c["S"] * R * P_s / (ms / 1e3)
    sample = ("%d of the %d permutations per step, all gene sets; C restatement of the reference loop "
              "(oracle/oracle.c), OpenMP over permutations" % (P_s, c["P"]))
    print(json.dumps({
        "impl": "reference", "metric": "gene-set x permutation ES/sec (null loop)", "value": val, "unit": "ES/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.config], "step": "bounded sample: " + sample},
        "cpu_baseline": {"value": val, "unit": "ES/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "ES/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "R is absent from this image (no R/Rscript/libR, no network): the reference package cannot run; "
                "this is the restated-reference CPU baseline (no R interpreter overhead, so a stronger baseline)",
    }))


def run_gpu(args):
    import torch
    import flexgsea_r_b200 as fg
    from flexgsea_r_b200 import _native as nat, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c, x, resp, icpt, off, idx = workload(args.config)
    G, n, S, P = c["G"], c["n"], c["S"], c["P"]
    if args.perms:
        P = args.perms
    _, score_name, es_name = KINDS[args.config]
    score_kind = nat.SCORE_LM if score_name == "lm" else nat.SCORE_S2N
    es_kind = {"wks": nat.ES_WEIGHTED_KS, "maxmean": nat.ES_MAXMEAN, "mean": nat.ES_MEAN}[es_name]
    nnz = int(off[-1])
    ctx = fg.Context(local)
    if world > 1:
        # the library's own NCCL communicator (include/flexgsea_b200.h): rank 0 makes the id, torch carries it
        idt = torch.zeros(nat.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nat.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)

    def upload(xh, rh, oh, ih):
        ctx.set_x(xh); ctx.set_gene_sets(oh, ih)
        if score_kind == nat.SCORE_LM:
            ctx.set_response_lm(rh, icpt)
        else:
            ctx.set_response_s2n(rh)

    upload(x, resp, off, idx)
    ctx.set_option("nperm_total", P)
    R = ctx.n_response(score_kind)
    ctx.set_profiling(True)
    seed = 20261019
    lo, hi = nat.shard_span(P, world, rank)      # strong scaling: the config's nperm is sharded
    sc_obs = ctx.score_observed(score_kind)
    es_obs = ctx.observed_es(es_kind, sc_obs, ragged=False)["es"]

    def step():
        """one pass of the hot path: this rank's span of the null loop, then the significance of
        every response over the sharded null (NCCL exchange inside the library when N > 1)"""
        ctx.perm_null(score_kind, es_kind, hi - lo, perm=None, seed=seed, perm_id0=lo, want_host=False)
        ms, ln = ctx.last_timings()
        coll = 0.0
        for r in range(R):
            sig = ctx.calc_sig_sharded(es_obs[:, r], P, r, True, True, False, exchange=args.exchange)
            coll += ctx.last_sig_timings()["collectives"]
        t = ctx.last_sig_timings()
        return ms, ln, t["step"], coll, sig

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.begin()
    l0 = ctx.launch_count()
    dev_ms, null_ms, coll_ms = 0.0, 0.0, 0.0
    stage_ms, stage_launch = {"perm": 0.0, "score": 0.0, "rank": 0.0, "es": 0.0}, None
    w0 = time.perf_counter()
    for _ in range(args.steps):
        ms, ln, step_ms, coll, sig = step()
        dev_ms += step_ms
        null_ms += ms["total"]
        coll_ms += coll
        for k in stage_ms:
            stage_ms[k] += ms[k]
        stage_launch = ln
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, wall_ms, null_ms, coll_ms], dtype=torch.float64, device="cuda")
    per_rank = None
    if dist is not None:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)     # what every rank saw: the spread is the wait inside the collectives
        per_rank = {"step_ms": [float(p[0]) / args.steps for p in parts],
                    "null_loop_ms": [float(p[2]) / args.steps for p in parts],
                    "collective_ms": [float(p[3]) / args.steps for p in parts]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, null_ms, coll_ms = (float(v) for v in t)
    lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(lt)
    launches = int(lt[0])
    units = float(S) * R * P * args.steps
    value = units / (dev_ms / 1e3)
    assert np.isfinite(sig["fdr"]).all() and (sig["p"] > 0).all()

    # ---- end to end through the public C-ABI calls of one flexgsea() run --------
    perm_h = synth.perm_matrix(n, P, seed=4) if P * n <= 4e7 else None   # C5: drawn on the device
    xp, _k1 = pinned(x); rp, _k2 = pinned(resp)
    pp = None
    if perm_h is not None:
        pp, _k3 = pinned(np.asfortranarray(perm_h[:, lo:hi]))
    offp, _k4 = pinned(off); idxp, _k5 = pinned(idx)

    def e2e_step():
        upload(xp, rp, offp, idxp)
        sc = ctx.score_observed(score_kind)
        obs = ctx.observed_es(es_kind, sc, ragged=False)  # default flexgsea(): return_values = character()
        ctx.perm_null(score_kind, es_kind, hi - lo, perm=pp, seed=seed, perm_id0=lo, want_host=False)
        return [ctx.calc_sig_sharded(obs["es"][:, r], P, r, True, True, False, exchange=args.exchange)
                for r in range(R)]

    e2e_step()
    barrier()
    e0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 3))
    for _ in range(e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - e0) / e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    # bytes over PCIe per step, all ranks: uploads are replicated, the permutation matrix is sharded
    h2d = world * (x.nbytes + resp.nbytes + 4 * (S + 1 + nnz) + 8 * G * R + 8 * S * R) + (4 * n * P if pp is not None else 0)
    d2h = world * (8 * G * R + S * R * (8 * 3 + 4 * 2) + R * (S * (6 * 8 + 4 * 8) + 32))

    if rank != 0:
        if dist is not None:
            ctx.comm_destroy()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------
    peak, peak_src = peaks()
    Gp = (G + 7) & ~7
    dom = max(("score", "rank", "es"), key=lambda k: stage_ms[k])
    wks = es_kind == nat.ES_WEIGHTED_KS
    alg_bytes_col = {  # algorithmic HBM bytes per score column (DESIGN.md "Kernels")
        "score": 8 * G,                 # score write (x is L2 resident, 8*n*G once)
        "rank": 8 * G + 10 * Gp,        # score read + rank (u16) and rank-ordered weight (f64) write
        "es": (10 * Gp if wks else 8 * G) + 8 * S,
    }
    cols = (hi - lo) * R   # rank 0's columns (the stage times are rank 0's)
    roof = {}
    for k in ("score", "rank", "es"):
        nl = max(1, stage_launch[k])
        t_ms = stage_ms[k] / args.steps
        once = 4 * (nnz + S + 1) if k == "es" else (8 * n * G if k == "score" else 0)
        b = cols * alg_bytes_col[k] + once * nl
        roof[k] = {"ms_per_step": t_ms, "launches_per_step": nl, "alg_bytes_per_step": b,
                   "achieved_gbs": b / (t_ms / 1e3) / 1e9 if t_ms > 0 else None}
    dl = roof[dom]
    per_launch_bytes = dl["alg_bytes_per_step"] / dl["launches_per_step"]
    per_launch_ms = dl["ms_per_step"] / dl["launches_per_step"]
    kname = {"score": "lm_gemm_kernel" if score_kind == nat.SCORE_LM else "s2n_mma_kernel",
             "rank": "sort_rank32_kernel" if G <= 29696 else "sort_rank_big_kernel",
             "es": "es_wks_kernel" if wks else "es_mean_smem_kernel"}[dom]
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": dl["achieved_gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": dl["achieved_gbs"] / peak, "traffic": None,
                "alg_bytes_per_launch": per_launch_bytes, "launch_ms": per_launch_ms,
                "note": "stage time from CUDA events on the library's stream inside the timed region"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:  # DRAM bytes per launch from the committed ncu --set full capture (per column x columns per launch)
            t = json.load(open(tr)).get(roofline["kernel"])
            if t and t.get("config", "C2") == KINDS[args.config][0]:
                roofline["traffic"] = t["dram_bytes_per_column"] * cols / dl["launches_per_step"]
                roofline["traffic_source"] = t["source"]
                if t.get("pipes"):  # what actually binds the kernel (ncu, same capture): not DRAM
                    roofline["pipes_pct_of_peak"] = t["pipes"]
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N == 1 only) ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as orc
        cores = orc.use_all_cores()
        sk, ek = oracle_kinds(orc, args.config)
        P_s = cpu_sample_perms(args.config, cores)
        ps = synth.perm_matrix(n, P_s, seed=4)
        orc.perm_null(x, sk, resp, icpt, ps[:, :max(1, min(cores, P_s))], off, idx, ek)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0 or reps < 1:
            rc, _ = orc.perm_null(x, sk, resp, icpt, ps, off, idx, ek)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": S * R * P_s / dt, "unit": "ES/s", "cores": cores, "kind": "port",
               "sample": "%d of the %d permutations, all %d gene sets, C restatement (oracle/oracle.c) with "
                         "OpenMP over permutations" % (P_s, P, S)}

    out = {
        "metric": "gene-set x permutation ES/sec (null loop + significance)", "value": value, "unit": "ES/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.config], "nperm": P, "nperm_per_gpu": hi - lo, "genes": G,
                   "samples": n, "gene_sets": S, "nnz": nnz, "responses": R,
                   "seeds": {"sets": 1, "x": 2, "philox": seed},
                   "permutations": "device Philox4x32-10 keyed by the global permutation id (value); host int32 "
                                   "matrix uploaded (e2e)" + ("" if pp is not None else " -- drawn on the device for this size"),
                   "sharding": "rank r takes permutations [r*P/N, (r+1)*P/N) of the config's nperm (strong scaling); "
                               "significance exchange: " + ("all-gather of the null shards" if args.exchange else
                               "all-gather of per-set double-double sums + int64 all-reduce of the counts")
                               + " over NCCL inside the library",
                   "l2": "inputs larger than L2: each block streams %d MB of scores/ranks/weights (L2 is 126 MB)"
                         % int(min((hi - lo) * R, 4096) * (18 * G) / 1e6),
                   "score_mode": ("FP64 tensor-core GEMM against the permuted pseudo-inverse rows" if score_kind == nat.SCORE_LM else
                                  "FP64 tensor-core contraction + rank certification (near-tied / cancelled entries "
                                  "recomputed in the reference operation order)")},
        "timed_region": "null loop of the rank's span + flexgsea_calc_sig of every response over the sharded null; "
                        "device time from the first kernel of the null loop to the last of the significance "
                        "(CUDA events on the library's stream), max over ranks",
        "null_loop_ms_per_step": null_ms / args.steps,
        "collective_ms_per_step": coll_ms / args.steps,
        "per_rank": per_rank,
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": {"value": float(S) * R * P / e2e_s, "unit": "ES/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                "what": "set_x + set_gene_sets + set_response + observed scores/ES(+extras) + null loop "
                        "(host permutation matrix) + calc_sig" + (" (sharded: NCCL exchange)" if world > 1 else "") + "; pinned host buffers; "
                        "result table read back on every rank"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "stages": roof,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))
    if dist is not None:
        ctx.comm_destroy()
        dist.destroy_process_group()


def run_gpu_group(args):
    """`python bench.py --gpus N` WITHOUT torchrun: one process, N GPUs through the library's group of
    contexts (fgs_group_*: one host thread per GPU, ncclCommInitAll) -- the path R's `parallel = N`
    binds.  Same step (sharded null loop + significance with the NCCL exchange), same fields."""
    import torch  # first: the library binds whichever libnccl.so.2 the process already holds (torch's, here)
    import flexgsea_r_b200 as fg
    from flexgsea_r_b200 import _native as nat, synth
    N = args.gpus
    c, x, resp, icpt, off, idx = workload(args.config)
    G, n, S, P = c["G"], c["n"], c["S"], c["P"]
    if args.perms:
        P = args.perms
    _, score_name, es_name = KINDS[args.config]
    score_kind = nat.SCORE_LM if score_name == "lm" else nat.SCORE_S2N
    es_kind = {"wks": nat.ES_WEIGHTED_KS, "maxmean": nat.ES_MAXMEAN, "mean": nat.ES_MEAN}[es_name]
    grp = nat.Group(N)

    def upload(xh, rh, oh, ih):
        grp.set_x(xh); grp.set_gene_sets(oh, ih)
        if score_kind == nat.SCORE_LM:
            grp.set_response_lm(rh, icpt)
        else:
            grp.set_response_s2n(rh)

    upload(x, resp, off, idx)
    R = grp.n_response(score_kind)
    seed = 20261019
    sc_obs = grp.score_observed(score_kind)
    es_obs = grp.observed_es(es_kind, sc_obs, ragged=False)["es"]

    def step():
        grp.perm_null(score_kind, es_kind, P, perm=None, seed=seed, want_host=False)
        coll = 0.0
        for r in range(R):
            sig = grp.calc_sig(es_obs[:, r], r, True, True, False, exchange=args.exchange)
            coll += grp.last_timings()["collectives"]
        t = grp.last_timings()
        return t["step"], t["null"], coll, sig

    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(args.warmup):
        step()
    sampler.begin()
    l0 = sum(cx.launch_count() for cx in grp.ctx)
    dev_ms = null_ms = coll_ms = 0.0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        st, nl, cl, sig = step()
        dev_ms += st; null_ms += nl; coll_ms += cl
    wall_ms = 1e3 * (time.perf_counter() - w0)
    launches = sum(cx.launch_count() for cx in grp.ctx) - l0
    clocks = sampler.stop()
    assert np.isfinite(sig["fdr"]).all()
    perm_h = synth.perm_matrix(n, P, seed=4) if P * n <= 4e7 else None
    xp, _k1 = pinned(x); rp, _k2 = pinned(resp); offp, _k4 = pinned(off); idxp, _k5 = pinned(idx)
    pp = pinned(perm_h)[0] if perm_h is not None else None

    def e2e_step():
        upload(xp, rp, offp, idxp)
        sc = grp.score_observed(score_kind)
        obs = grp.observed_es(es_kind, sc, ragged=False)
        grp.perm_null(score_kind, es_kind, P, perm=pp, seed=seed, want_host=False)
        return [grp.calc_sig(obs["es"][:, r], r, True, True, False, exchange=args.exchange) for r in range(R)]

    e2e_step()
    e0 = time.perf_counter()
    e_steps = max(1, min(args.steps, 3))
    for _ in range(e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - e0) / e_steps
    nnz = int(off[-1])
    h2d = N * (x.nbytes + resp.nbytes + 4 * (S + 1 + nnz) + 8 * G * R + 8 * S * R) + (4 * n * P if pp is not None else 0)
    d2h = 8 * G * R + S * R * (8 * 3 + 4 * 2) + R * (S * (6 * 8 + 4 * 8) + 32)
    print(json.dumps({
        "metric": "gene-set x permutation ES/sec (null loop + significance)", "value": float(S) * R * P * args.steps / (dev_ms / 1e3),
        "unit": "ES/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.config], "nperm": P, "genes": G, "samples": n, "gene_sets": S,
                   "nnz": nnz, "responses": R, "seeds": {"sets": 1, "x": 2, "philox": seed},
                   "sharding": "ONE process, %d GPUs through fgs_group_* (one host thread per GPU, ncclCommInitAll): "
                               "GPU g takes permutations [g*P/N, (g+1)*P/N); NCCL exchange inside the library" % N,
                   "l2": "inputs larger than L2: every block streams far more than 126 MB of scores/ranks/weights"},
        "timed_region": "max over the GPUs of (first kernel of the null loop -> last of the significance), CUDA events",
        "null_loop_ms_per_step": null_ms / args.steps, "collective_ms_per_step": coll_ms / args.steps,
        "wall_ms_per_step": wall_ms / args.steps,
        "e2e": {"value": float(S) * R * P / e2e_s, "unit": "ES/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                "what": "the fgs_group_* calls of one flexgsea(parallel = N): uploads to every GPU, observed pass, "
                        "sharded null loop (host permutation matrix), significance with the NCCL exchange"},
        "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None, "clocks": clocks,
    }))
    grp.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="C2", choices=list(WORKLOAD_TEXT))
    ap.add_argument("--perms", type=int, default=0, help="override the config's nperm (debug)")
    ap.add_argument("--ref-perms", type=int, default=0, help="permutations per reference step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", type=int, default=0, choices=[0, 1],
                    help="significance exchange at N > 1: 0 = counts (default), 1 = all-gather of the null")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "native" and (args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1):
        # The exchange is a few hundred KB: NVLink-SHARP (NVLS) multicast buys nothing, and with it enabled
        # every nvidia-smi query of the clock sampler stalled the sampled rank for milliseconds (two ranks on an
        # 8-GPU box: 3.7 ms per step of waiting in the collectives; 0.35 ms with NVLS off, 0.14 ms without the
        # sampler -- profiles/README.md).  Set before any NCCL communicator exists.
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        run_gpu_group(args)      # no torchrun: one process drives the N GPUs (what R's parallel = N does)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
