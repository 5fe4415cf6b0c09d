#!/usr/bin/env python
"""bench.py -- evaluated users/sec of the full-ranking evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2]

One "step" = one full pass of the hot path (operand split, fused score+mask+top-K, partial merge,
metrics, column sums, and for N > 1 the metric all-reduce) over the workload: BASELINE.json
configs[1], LightGCN on Gowalla-shape synthetic data (29,858 users x 40,981 items, d=64, 810,128
masked train interactions, top-[20,50] Precision/Recall/NDCG).  N > 1: every rank evaluates its own
29,858-user slice against the replicated item table (user-sharded, weak scaling) and only the
metric sums are all-reduced over NCCL.

value : inputs resident in HBM, CUDA-event time of the steps on the launching stream (L2 flushed
        between steps, outside the event pairs), max over ranks.
e2e   : the same metric through RankingEvaluator.evaluate(model) with HOST (pinned) embedding
        tables: H2D of the tables and D2H of the metric sums inside the timed region.
roofline : the scoring kernel alone (CUDA events recorded around that launch inside the library).
cpu_baseline / --impl reference : the UNMODIFIED reference evaluator (oracle/_ref, compiled from
        /root/reference) with a torch-CPU `predict`, all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "evaluated users/sec (full-rank top-K)"
UNIT = "users/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), src="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")  # B200_PROFILING.md


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": "warm-up + timed steps + e2e steps"}


C4_RANKS = 8  # c4 is quoted user-sharded over 8 B200: one rank's share is 125,000 of the 1,000,000 users


def rank_config(name, users=None):
    """The workload one rank holds.  c1-c3: the whole config (N > 1 gives every rank its own slice of that size).
    c4 (1M x 1M): one GPU's share of the 8-way user sharding -- 125,000 users against the full 1M-item table, interactions
    scaled with the users (50 train / 10 test items per user) -- so `--gpus 8` is exactly c4 and `--gpus 1` one eighth of it."""
    from skrec_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    if name == "c4" and users is None:
        users = cfg["users"] // C4_RANKS
    if users is not None and users != cfg["users"]:
        f = float(users) / cfg["users"]
        cfg["nnz_train"] = max(users, int(round(cfg["nnz_train"] * f)))
        cfg["nnz_test"] = max(users, int(round(cfg["nnz_test"] * f)))
        cfg["name"] = "%s, %d-user slice" % (cfg["name"], users)
        cfg["users"] = int(users)
    return cfg


def reference_users_per_s(data, cfg, users, cores, batch_size=256, repeats=1):
    """The unmodified reference RankingEvaluator.evaluate (oracle/_ref) on `users`; best of repeats."""
    import torch
    import oracle
    torch.set_num_threads(cores)
    ue, ie = torch.from_numpy(data["user_emb"]), torch.from_numpy(data["item_emb"])
    bias = None if data["bias"] is None else torch.from_numpy(data["bias"])

    class Model(object):  # the `predict` of a dot-product recommender (LightGCN.py:102-107,214-216)
        def predict(self, us):
            s = torch.matmul(ue[torch.as_tensor(us)], ie.T)
            if bias is not None:
                s = s + bias
            return s.cpu().detach().numpy()

    ev = oracle.RefRankingEvaluator(data["train"], data["test"], metric=cfg["metric"], top_k=cfg["top_k"],
                                    batch_size=batch_size, num_thread=cores)
    best, rep = None, None
    for _ in range(repeats):
        t = time.perf_counter()
        rep = ev.evaluate(Model(), test_users=users)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return len(users) / best, best, rep


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from skrec_b200 import synth
    if not oracle.ref_python_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) not present"}))
        return 0
    import warnings
    warnings.filterwarnings("ignore")
    # c4: a 2,048-user slice against the full 1M-item table (a full pass would stream 4 TB of scores)
    cfg = rank_config(args.config, users=2048 if args.config == "c4" else args.users_per_gpu)
    if args.config == "c4":
        cfg["item_seed"] = cfg["seed"] + 7
    data = synth.make(**cfg)
    cores = os.cpu_count() or 1
    U = data["users"]
    rate, _, _ = reference_users_per_s(data, cfg, list(range(min(U, 1024))), cores)
    total_steps = args.steps + args.warmup
    n = int(min(U, max(512, rate * 150.0 / max(total_steps, 1))))
    users = list(range(n))
    for _ in range(args.warmup):
        reference_users_per_s(data, cfg, users, cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        reference_users_per_s(data, cfg, users, cores)
    dt = (time.perf_counter() - t) / max(args.steps, 1)
    value = n / dt
    sample = "%d of %d users per step (contiguous from user 0), batch 256, %d threads" % (n, U, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.config, cfg["name"]), "users": U, "items": data["items"], "d": data["d"],
                       "top_k": cfg["top_k"], "metrics": cfg["metric"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as td
    from skrec_b200 import RankingEvaluator, _native, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)

    cfg = rank_config(args.config, users=args.users_per_gpu)
    if world > 1 or args.config == "c4":  # each rank: its own user slice of one replicated catalogue
        cfg["item_seed"] = cfg["seed"] + 7
        cfg["seed"] = cfg["seed"] + 1000 * rank
    data = synth.make(device=dev, **cfg)
    U, I, d = data["users"], data["items"], data["d"]
    K = max(cfg["top_k"])
    ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
    MK = len(ids) * K

    # ---- device-resident arm ------------------------------------------------------------------
    ctx = _native.Context(local)
    ctx.set_train_csr(data["train_indptr"], data["train_indices"], I)
    ctx.set_test_csr(data["test_indptr"], data["test_indices"], I)
    ue = torch.from_numpy(data["user_emb"]).to(dev)
    ie = torch.from_numpy(data["item_emb"]).to(dev)
    bias = None if data["bias"] is None else torch.from_numpy(data["bias"]).to(dev)
    sums = torch.zeros(MK + 1, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # [column sums | user count] is what the ranks exchange: the accumulator starts every step as a device-side copy of
    # [0 ... 0, U] (assigning a Python scalar to sums[MK] costs a blocking host-to-device copy: measured 44 us per step)
    sums_init = torch.zeros(MK + 1, dtype=torch.float64, device=dev)
    sums_init[MK] = float(U)

    def step():
        sums.copy_(sums_init)
        ctx.eval_fused(ue, ie, bias, 0, ids, K, precision=args.precision, sums=sums[:MK])
        if world > 1:
            td.all_reduce(sums)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load0 = time.time()
    # >= W steps and about 1 s of load so the clocks ramp.  The count must be the same on every rank (each step
    # holds an all-reduce), so it is fixed up front from the step time rank 0 measures, not decided by a clock.
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()
    t_probe = time.time()
    for _ in range(20):
        step()
    torch.cuda.synchronize()
    extra = torch.tensor([int(min(5000, max(0.0, 1.0 / max((time.time() - t_probe) / 20, 1e-5))))], dtype=torch.int64, device=dev)
    if world > 1:
        td.broadcast(extra, 0)
    n_extra = int(extra.item())
    for i in range(n_extra):
        step()
        if i % 16 == 15:
            torch.cuda.synchronize()
    n_warm += 20 + n_extra
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ctx.set_option("event_ring", args.steps)
    launches0 = ctx.launch_count
    for a, b in ev:
        flush.zero_()
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0  # this library's kernels only (torch adds the L2 flush and sums.zero_() per step, NCCL the all-reduce)
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = [ctx.fused_kernel_ms(i) for i in range(args.steps)]
    prepass_ms = [ctx.fused_prepass_ms(i) for i in range(args.steps)]
    plan = ctx.fused_stats()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * U / (ms_per_step * 1e-3)
    means = (sums[:MK] / float(world * U)).cpu().numpy().reshape(len(ids), K)[:, np.array(cfg["top_k"]) - 1].ravel()

    # ---- end to end through the public API, host (pinned) tables ---------------------------------
    ue_h = torch.from_numpy(data["user_emb"]).pin_memory()
    ie_h = torch.from_numpy(data["item_emb"]).pin_memory()
    b_h = None if data["bias"] is None else torch.from_numpy(data["bias"]).pin_memory()

    class HostModel(object):
        def predict(self, users):  # reference protocol; unused when eval_embeddings exists
            raise NotImplementedError

        def eval_embeddings(self, users):
            return (ue_h if len(users) == U else ue_h[torch.as_tensor(users)]), ie_h, b_h

    evaluator = RankingEvaluator(data["train"], data["test"], metric=cfg["metric"], top_k=cfg["top_k"], device=local,
                                 precision=args.precision, shard_users=False)
    model = HostModel()
    for _ in range(3):
        rep = evaluator.evaluate(model)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    e2e_steps = max(3, min(args.steps, 50))
    # host-side wall clock picks up whatever else the box's CPUs are doing: three blocks of e2e_steps, the best block counts
    # (the report is per rank; ranks run concurrently and are timed as max below)
    e2e_blocks = []
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rep = evaluator.evaluate(model)
        torch.cuda.synchronize()
        e2e_blocks.append((time.perf_counter() - t0) / e2e_steps)
    e2e_s = min(e2e_blocks)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_s = float(t.item())
    t_load1 = time.time()
    e2e = {"value": world * U / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": int(4 * (U * d + I * d + (I if b_h is not None else 0))), "d2h_bytes_per_step": int(8 * (MK + 1)),
           "api": "RankingEvaluator.evaluate(model) with pinned host embedding tables",
           "timing": "wall clock, best of 3 blocks of %d evaluate() calls; blocks (ms/step): %s" % (e2e_steps, ", ".join("%.3f" % (x * 1e3) for x in e2e_blocks))}

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return 0
    clocks = sampler.stop(t_load0, t_load1)

    # ---- roofline of the scoring kernel -----------------------------------------------------------
    pk = peaks()
    k_ms = float(np.mean(kernel_ms))
    flops = 2.0 * U * I * d
    achieved = flops / (k_ms * 1e-3) / 1e12
    passes = {"tcgen05_3xtf32": 3, "tcgen05_1xtf32": 1, "tcgen05_tf32r": 1}.get(ctx.last_fused_kernel, 1)
    if ctx.last_fused_kernel.startswith("tcgen05"):
        peak = pk["bf16"] / 2.0
        peak_note = "TF32 dense = 1/2 of the %s cuBLAS bf16 burst peak (%.1f TFLOP/s) in MEASURED_PEAKS.json" % (pk["src"], pk["bf16"])
    else:
        peak = 2 * 128 * 148 * 1.965e9 / 1e12  # FP32 FMA pipe: 128 lanes x 148 SMs x max clock
        peak_note = "FP32 FMA pipe nominal (128 FMA/clk/SM x 148 SMs x 1965 MHz)"
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full): "<kernel>" is the c2 capture,
        # "<kernel>@<config>" the others; configs without a capture report null
        tj = json.load(open(tp))
        traffic = tj.get("%s@%s" % (ctx.last_fused_kernel, args.config), tj.get(ctx.last_fused_kernel) if args.config == "c2" else None)
        if args.users_per_gpu is not None:
            traffic = None
    cublas_tf32 = None  # cuBLAS TF32 8192^3 on a B200 of this pool (tools/tf32_peak.py): a cross-check of the derived peak
    cp = os.path.join(ROOT, "profiles", "r1_run9_tf32_peak.json")
    if os.path.exists(cp):
        cublas_tf32 = json.load(open(cp))["cublas_8192_cubed"]["tf32"]["best_tflops"]
    roofline = {"bound": "tensor", "kernel": ctx.last_fused_kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                "prepass_kernel_ms": float(np.mean(prepass_ms)), "plan": plan,
                "algorithmic_flops_per_launch": flops, "mma_passes": passes, "tensor_pipe_utilisation_est": passes * achieved / peak,
                "peak_note": peak_note, "cublas_tf32_8192_tflops": cublas_tf32}

    # ---- the reference on this box's host cores, same workload -----------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        import warnings
        warnings.filterwarnings("ignore")
        cores = os.cpu_count() or 1
        if oracle.ref_python_available():
            n = U if U <= 60000 else 2048
            rate, secs, ref_rep = reference_users_per_s(data, cfg, list(range(n)), cores, repeats=2)
            ref_vals = np.array(list(ref_rep.values()), np.float32)
            gpu_vals = means
            if n != U:  # the same users through the GPU evaluator, so the sample is a parity check too
                gpu_vals = np.array(list(evaluator.evaluate(model, test_users=list(range(n))).values()), np.float32)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": "%d of %d users, 1 pass, best of 2 (%.2f s), batch 256, unmodified reference RankingEvaluator + torch CPU predict" % (n, U, secs),
                   "max_abs_metric_diff_vs_gpu": float(np.max(np.abs(ref_vals - gpu_vals)))}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref not present"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32x3 (fp32-grade)" if passes == 3 else ("tf32 candidates + f32 re-scoring (fp32-exact)" if ctx.last_fused_kernel == "tcgen05_tf32r"
                                                                else ("tf32" if ctx.last_fused_kernel.startswith("tcgen05") else "f32")),
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.config, cfg["name"]), "users_per_gpu": U, "items": I, "d": d, "train_nnz": int(data["train_indptr"][-1]),
                       "top_k": cfg["top_k"], "metrics": cfg["metric"], "parallelism": "user-sharded x%d, item table replicated, metric-sum all-reduce" % world,
                       "l2": "256 MB buffer written between steps, outside the event pairs", "precision": args.precision},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "metrics_at_top_k": {n_: float(v) for n_, v in zip(evaluator.metrics_list, means)}}
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 200 (c1-c3), 20 (c4: a step is ~70 ms)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3a", "c3b", "c4"])
    ap.add_argument("--users-per-gpu", type=int, default=None, help="evaluate a slice of this many users per rank (interactions scaled with it)")
    ap.add_argument("--precision", default="auto", choices=["auto", "3xtf32", "fp32", "1xtf32", "tf32r"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.config == "c4" else 200
    args.steps = max(1, args.steps)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
