#!/usr/bin/env python
"""bench.py -- evaluated users/sec of the full-ranking evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c4] [--scaling strong|weak]
                    [--path fused|predict]

Default workload = the configuration BASELINE.json's target is quoted on: **c4**, the MultVAE/SelfCF-style dense
scorer on 1,000,000 users x 1,000,000 items, d = 128, bias, ~50 masked train items per user, top-[10,20,50,100],
all five metrics.  One "step" = one full evaluation of ALL 10^6 users: item-table split, sampled thresholds, fused
score + mask + candidate pass on tcgen05, exact re-scoring + top-K + metrics, column sums and -- for N > 1 -- the
all-reduce of [sums | user count].  **Strong scaling**: with N ranks (torchrun, one per GPU) the product's own
`RankingEvaluator(..., shard_users=True)` gives every rank a contiguous 1/N of the users against the replicated item
table; the collective is inside `evaluate`.  c5 (2M users x 10M items) runs item-sharded (`shard="items"`:
per-rank top-K, all-gather of the rank keys overlapped with the next chunk, merge) the same way.

value    : whole-job users/s, tables resident in HBM, CUDA-event time of `RankingEvaluator.evaluate_device(model)`
           (= `evaluate` minus its final 4 KB device-to-host copy) on the launching stream, L2 flushed between steps
           (outside the event pairs), max over ranks.
e2e      : the same through `RankingEvaluator.evaluate(model)` with HOST (pinned) embedding tables: H2D of the tables
           and D2H of the result inside the timed region, wall clock, max over ranks.
roofline : the scoring kernel alone (CUDA events recorded around that launch inside the library, summed over the
           row chunks of a step).
cpu_baseline / --impl reference : the UNMODIFIED reference evaluator (oracle/_ref, compiled from /root/reference)
           with a torch-CPU `predict`, all host cores, on a bounded sample of the same workload.
also     : the round-1 headline workload (c2, weak-scaled: N x 29,858 users) measured the same way, compact.
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
LARGE = ("c4", "c5")          # generated on the GPU (synth.make_large), reference arm on a 2,048-user sample
STRONG = ("c3a", "c3b", "c4", "c5")
SAMPLE_USERS = 2048


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

    def stop(self, windows):
        """windows: [(t0, t1)] wall-clock intervals during which the GPU was under this bench's load"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not any(t0 <= t <= t1 + 0.1 for t0, t1 in windows):
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
                "window": "warm-up + timed steps + e2e steps of the headline workload"}


def workload_config(name, scaling):
    """The `config` object both arms print: only what defines the workload."""
    from skrec_b200 import synth
    cfg = synth.CONFIGS[name]
    return {"workload": "%s: %s" % (name, cfg["name"]), "users": cfg["users"], "items": cfg["items"], "d": cfg["d"],
            "bias": bool(cfg["bias"]), "train_nnz_nominal": cfg["nnz_train"], "test_nnz_nominal": cfg["nnz_test"],
            "top_k": cfg["top_k"], "metrics": cfg["metric"], "scaling": scaling}


def sample_config(name, users=SAMPLE_USERS):
    """A `users`-row slice of a large config against its full catalogue, interactions scaled with the users:
    what the CPU reference can finish (a full pass of c4 streams 4 TB of scores)."""
    from skrec_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    f = float(users) / cfg["users"]
    cfg["nnz_train"] = max(users, int(round(cfg["nnz_train"] * f)))
    cfg["nnz_test"] = max(users, int(round(cfg["nnz_test"] * f)))
    cfg["users"] = int(users)
    cfg["item_seed"] = cfg["seed"] + 7
    return cfg


# ---- the reference on the host cores ----------------------------------------------------------------------------
def reference_users_per_s(data, cfg, users, cores, batch_size=256, repeats=1):
    """The unmodified reference RankingEvaluator.evaluate (oracle/_ref) on `users`; best of repeats."""
    import torch
    import oracle
    torch.set_num_threads(cores)
    ue, ie = torch.as_tensor(data["user_emb"]), torch.as_tensor(data["item_emb"])
    bias = None if data["bias"] is None else torch.as_tensor(data["bias"])

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
    large = args.config in LARGE
    cfg = sample_config(args.config) if large else dict(synth.CONFIGS[args.config])
    data = synth.make(**cfg)
    cores = os.cpu_count() or 1
    U = data["users"]
    rate, _, _ = reference_users_per_s(data, cfg, list(range(min(U, 512 if large else 1024))), cores)
    total_steps = args.steps + args.warmup
    n = int(min(U, max(256, rate * 150.0 / max(total_steps, 1))))
    users = list(range(n))
    for _ in range(args.warmup):
        reference_users_per_s(data, cfg, users, cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        reference_users_per_s(data, cfg, users, cores)
    dt = (time.perf_counter() - t) / max(args.steps, 1)
    value = n / dt
    sample = "%d users per step (contiguous from user 0%s), batch 256, %d threads, unmodified reference RankingEvaluator + torch CPU predict" % (
        n, " of a %d-user slice against the full %d-item catalogue" % (U, data["items"]) if large else " of %d" % U, cores)
    scaling = args.scaling or ("strong" if args.config in STRONG else "weak")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, scaling),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---- our arm ----------------------------------------------------------------------------------------------------
class Problem(object):
    """One global evaluation problem, identical on every rank: host CSRs of all users, the item table (and bias)
    on the device, the user table on the device and -- for the end-to-end arm -- in pinned host memory."""

    def __init__(self, name, scaling, world, rank, dev, norms="iid"):
        import torch
        from skrec_b200 import synth
        cfg = dict(synth.CONFIGS[name], norms=norms)
        self.name, self.cfg, self.scaling = name, cfg, scaling
        t0 = time.time()
        if name in LARGE:
            d = synth.make_large(device=dev, item_seed=cfg["seed"] + 7, **cfg)
            self.train = (d["train_indptr"], d["train_indices"])
            self.test = (d["test_indptr"], d["test_indices"])
            self.ue, self.ie, self.bias = d["user_emb"], d["item_emb"], d["bias"]
        elif scaling == "weak" and world > 1:
            # N slices of `users` users each, one shared catalogue: the global problem grows with N
            parts = [synth.make(device=dev, **dict(cfg, seed=cfg["seed"] + 1000 * r, item_seed=cfg["seed"] + 7)) for r in range(world)]
            def cat(ptr_key, idx_key):
                ptr, off = [np.zeros(1, np.int64)], 0
                for p in parts:
                    ptr.append(p[ptr_key][1:] + off)
                    off += int(p[ptr_key][-1])
                return np.concatenate(ptr), np.concatenate([p[idx_key] for p in parts])
            self.train, self.test = cat("train_indptr", "train_indices"), cat("test_indptr", "test_indices")
            self.ue = torch.from_numpy(np.concatenate([p["user_emb"] for p in parts])).to(dev)
            self.ie = torch.from_numpy(parts[0]["item_emb"]).to(dev)
            self.bias = None if parts[0]["bias"] is None else torch.from_numpy(parts[0]["bias"]).to(dev)
        else:
            d = synth.make(device=dev, **cfg)
            self.train = (d["train_indptr"], d["train_indices"])
            self.test = (d["test_indptr"], d["test_indices"])
            self.ue = torch.from_numpy(d["user_emb"]).to(dev)
            self.ie = torch.from_numpy(d["item_emb"]).to(dev)
            self.bias = None if d["bias"] is None else torch.from_numpy(d["bias"]).to(dev)
            self.dicts = (d["train"], d["test"])
        self.U, self.I, self.d = int(self.ue.shape[0]), int(self.ie.shape[0]), int(self.ie.shape[1])
        self.gen_s = time.time() - t0
        self.host = None

    def pin_host(self, rows=None):
        """Pinned host copies for the end-to-end arm (`rows`: only this slice of the user table is ever asked for)."""
        lo, hi = rows if rows is not None else (0, self.U)
        self.host = dict(lo=lo, hi=hi, ue=self.ue[lo:hi].cpu().pin_memory(), ie=self.ie.cpu().pin_memory(),
                         bias=None if self.bias is None else self.bias.cpu().pin_memory())


def _span(users):
    """(first, count) when `users` is a run of consecutive ids (what a contiguous shard of all test users is), else None"""
    n = len(users)
    if n and int(users[-1]) - int(users[0]) + 1 == n:
        return int(users[0]), n
    return None


class DeviceModel(object):
    """A trained dot-product recommender whose tables live on the GPU.  `eval_embeddings` is the optional protocol of
    the fused path (operands of `predict`: MultVAE.py:138-141, SelfCF.py:235-241, LightGCN.py:102-107)."""

    def __init__(self, prob, dev):
        self.p, self.dev = prob, dev

    def predict(self, users):  # reference protocol (base.py:73-74): host float32 [B, I]
        import torch
        s = self.p.ue[torch.as_tensor(np.asarray(users, dtype=np.int64), device=self.dev)] @ self.p.ie.T
        if self.p.bias is not None:
            s = s + self.p.bias
        return s.cpu().detach().numpy()

    def _rows(self, table, users, base=0):
        import torch
        sp = _span(users)
        if sp is not None:
            return table[sp[0] - base:sp[0] - base + sp[1]]
        return table[torch.as_tensor(np.asarray(users, dtype=np.int64) - base, device=table.device)]

    def eval_embeddings(self, users, item_shard=None):
        from skrec_b200 import dist
        uv = self._rows(self.p.ue, users)
        if item_shard is None:
            return uv, self.p.ie, self.p.bias
        lo, hi = dist.shard_range(self.p.I, item_shard[0], item_shard[1])
        return uv, self.p.ie[lo:hi], None if self.p.bias is None else self.p.bias[lo:hi], self.p.I


class HostModel(DeviceModel):
    """The same model with its tables in (pinned) host memory: every evaluate uploads what it needs."""

    def eval_embeddings(self, users, item_shard=None):
        from skrec_b200 import dist
        h = self.p.host
        uv = self._rows(h["ue"], users, base=h["lo"])
        if item_shard is None:
            return uv, h["ie"], h["bias"]
        lo, hi = dist.shard_range(self.p.I, item_shard[0], item_shard[1])
        return uv, h["ie"][lo:hi], None if h["bias"] is None else h["bias"][lo:hi], self.p.I


class PredictModel(object):
    """Only the reference protocol: `predict(users)` -> host float32 [B, I] (what an unmodified reference model gives:
    a GPU GEMM, then `.cpu().detach().numpy()`, LightGCN.py:214-216)."""

    def __init__(self, prob, dev):
        self._m = DeviceModel(prob, dev)

    def predict(self, users):
        return self._m.predict(users)


def measure(args, name, scaling, steps, warmup, e2e_cap, dev, rank, world, local, sampler_windows, headline):
    """Build the problem, time `steps` evaluations (device-resident and end to end) -> dict of results (every rank)."""
    import torch
    import torch.distributed as td
    from skrec_b200 import RankingEvaluator, dist

    prob = Problem(name, scaling, world, rank, dev, norms=args.norms if headline else "iid")
    cfg = prob.cfg
    K = max(cfg["top_k"])
    M = len(cfg["metric"])
    MK = M * K
    shard = "items" if name == "c5" else "users"
    evaluator = RankingEvaluator.from_csr(prob.train, prob.test, metric=cfg["metric"], top_k=cfg["top_k"], device=local,
                                          precision=args.precision, shard_users=True, shard=shard,
                                          batch_size=args.batch_size)
    predict_path = args.path == "predict"
    model = PredictModel(prob, dev) if predict_path else DeviceModel(prob, dev)
    U = prob.U
    lo, hi = dist.shard_range(U, rank, world) if (shard == "users") else (0, U)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        return evaluator.evaluate_device(model)

    t_prep = time.time()
    packed = step()  # builds the plan: CSR slices -> device (set-up, reported as prep_ms)
    torch.cuda.synchronize()
    prep_ms = (time.time() - t_prep) * 1e3
    ctx = evaluator.last_context()
    t_w0 = time.time()
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    # short steps: keep the GPU busy for about a second so the clocks ramp (same count on every rank: each step
    # holds a collective).  Reported separately; --warmup is honoured as given.
    per = (time.time() - t_w0) / max(warmup, 3)
    extra = torch.tensor([int(min(3000, max(0.0, (1.0 - (time.time() - t_w0)) / max(per, 1e-5))))], dtype=torch.int64, device=dev)
    if world > 1:
        td.broadcast(extra, 0)
    n_ramp = int(extra.item())
    for i in range(n_ramp):
        step()
        if i % 16 == 15:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if ctx is not None and not predict_path:
        ctx.set_option("event_ring", 8192)
    launches0 = ctx.launch_count if ctx is not None else 0
    for a, b in ev:
        flush.zero_()
        a.record()
        packed = step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    launches = (ctx.launch_count - launches0) if ctx is not None else 0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    res = dict(prob=prob, evaluator=evaluator, U=U, K=K, M=M, I=prob.I, d=prob.d, prep_ms=prep_ms, n_ramp=n_ramp,
               path=evaluator.last_stats.get("path"), gen_s=prob.gen_s, names=list(evaluator.metrics_list))
    fused = ctx is not None and not predict_path and ctx.last_fused_kernel != "none"
    if fused:
        st = ctx.fused_stats()
        per_step = max(1, st["timed_launches"] // steps)
        k_ms = [sum(ctx.fused_kernel_ms(s * per_step + c) for c in range(per_step)) for s in range(steps)]
        p_ms = [sum(ctx.fused_prepass_ms(s * per_step + c) for c in range(per_step)) for s in range(steps)]
        res.update(kernel_ms=float(np.mean(k_ms)), prepass_ms=float(np.mean(p_ms)), plan=st, chunks_per_step=per_step,
                   kernel=ctx.last_fused_kernel)
    t = torch.tensor([total_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
        td.all_reduce(t, op=td.ReduceOp.SUM)
        total_ms, launches = float(tmax[0].item()), float(t[1].item())
    res["ms_per_step"] = total_ms / steps
    res["value"] = U / (res["ms_per_step"] * 1e-3)
    res["launches"] = int(launches)
    host = packed.cpu().numpy()
    res["means"] = (host[:MK] / host[MK]).astype(np.float32).reshape(M, K)[:, np.array(cfg["top_k"]) - 1].ravel()
    res["users_counted"] = int(host[MK])

    # ---- end to end through the public API, host (pinned) tables -----------------------------------
    if predict_path:
        e2e_model = model  # `predict` returns host blocks: the value above already is end to end
    else:
        prob.pin_host((lo, hi))
        e2e_model = HostModel(prob, dev)
    for _ in range(2):
        rep = evaluator.evaluate(e2e_model)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    e2e_steps = max(2, min(steps, e2e_cap))
    blocks = []
    for _ in range(2 if e2e_steps * res["ms_per_step"] > 1500 else 3):
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rep = evaluator.evaluate(e2e_model)
        torch.cuda.synchronize()
        blocks.append((time.perf_counter() - t0) / e2e_steps)
    e2e_s = min(blocks)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    e2e_s = float(t.item())
    if predict_path:
        h2d = 4 * U * prob.I
    elif shard == "items":
        h2d = 4 * (world * U * prob.d + prob.I * prob.d + (prob.I if prob.bias is not None else 0))
    else:
        # every rank uploads its users; the item table once in total when sharded (1/world per rank, all-gathered over
        # NVLink: RankingEvaluator(upload="sharded"), the default), the bias vector whole
        tables = 1 if (world > 1 and evaluator.upload == "sharded") else world
        h2d = 4 * (U * prob.d + tables * prob.I * prob.d + world * (prob.I if prob.bias is not None else 0))
    res["e2e"] = {"value": U / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": int(h2d),
                  "d2h_bytes_per_step": int(8 * (MK + 1) * world),
                  "api": ("RankingEvaluator.evaluate(model), model.predict -> host score blocks" if predict_path else
                          "RankingEvaluator(shard_users=True%s).evaluate(model) with pinned host embedding tables%s"
                          % (", shard='items'" if shard == "items" else "",
                             " (item table: 1/%d uploaded per rank, all-gathered over NVLink)" % world
                             if (world > 1 and shard == "users" and evaluator.upload == "sharded") else ""))
                         + "; bytes summed over the %d rank(s)" % world,
                  "timing": "wall clock, best of %d blocks of %d evaluate() calls, max over ranks; blocks on rank %d (ms/step): %s"
                            % (len(blocks), e2e_steps, rank, ", ".join("%.3f" % (x * 1e3) for x in blocks))}
    res["report"] = rep
    if predict_path:
        # the HBM-bound kernel on a resident score block: 4 I bytes per user is all it has to read
        B = min(U, max(256, min(8192, (2 << 30) // (4 * prob.I))))
        blk = torch.from_numpy(model.predict(list(range(B)))).to(dev)
        sums = torch.zeros(MK, dtype=torch.float64, device=dev)
        for _ in range(3):
            ctx.eval_scores(blk, 0, evaluator.metrics, K, sums=sums)
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in pairs:
            flush.zero_()
            a.record()
            ctx.eval_scores(blk, 0, evaluator.metrics, K, sums=sums)
            b.record()
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
        res["scores_kernel"] = dict(rows=B, ms=ms, gbs=4.0 * B * prob.I / (ms * 1e-3) / 1e9)
    if headline:
        sampler_windows.append((t_w0, time.time()))
    return res


def run_ours(args):
    import torch
    import torch.distributed as td
    from skrec_b200 import RankingEvaluator, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=dev)

    name = args.config
    scaling = args.scaling or ("strong" if name in STRONG else "weak")
    sampler = ClockSampler(local)
    windows = []
    if rank == 0:
        sampler.start()
    e2e_cap = {"c4": 4, "c5": 2}.get(name, 50)
    r = measure(args, name, scaling, args.steps, args.warmup, e2e_cap, dev, rank, world, local, windows, True)
    prob, evaluator, U, K, M = r["prob"], r["evaluator"], r["U"], r["K"], r["M"]
    cfg = prob.cfg
    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- the reference on this box's host cores, same workload (N = 1 only) ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        import warnings
        warnings.filterwarnings("ignore")
        cores = os.cpu_count() or 1
        if oracle.ref_python_available():
            if name in LARGE:  # a slice the reference can finish; the GPU evaluator runs the same slice for the parity figure
                scfg = sample_config(name)
                sd = synth.make(device=dev, **scfg)
                n = sd["users"]
                rate, secs, ref_rep = reference_users_per_s(sd, scfg, list(range(n)), cores, repeats=1)
                sev = RankingEvaluator(sd["train"], sd["test"], metric=scfg["metric"], top_k=scfg["top_k"], device=local, precision=args.precision)
                gpu_vals = np.array(list(sev.evaluate(synth.EmbeddingModel(sd["user_emb"], sd["item_emb"], sd["bias"])).values()), np.float32)
                what = "%d-user slice against the full %d-item catalogue (same recipe, interactions scaled), 1 pass (%.2f s)" % (n, sd["items"], secs)
            else:
                n = U
                data = dict(user_emb=prob.ue.cpu().numpy(), item_emb=prob.ie.cpu().numpy(),
                            bias=None if prob.bias is None else prob.bias.cpu().numpy(), train=prob.dicts[0], test=prob.dicts[1])
                rate, secs, ref_rep = reference_users_per_s(data, cfg, list(range(n)), cores, repeats=2)
                gpu_vals = r["means"]
                what = "all %d users, 1 pass, best of 2 (%.2f s)" % (n, secs)
            ref_vals = np.array(list(ref_rep.values()), np.float32)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": what + ", batch 256, unmodified reference RankingEvaluator + torch CPU predict",
                   "max_abs_metric_diff_vs_gpu": float(np.max(np.abs(ref_vals - gpu_vals)))}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "oracle/_ref not present"}

    # ---- the round-1 headline workload next to it -------------------------------------------------------
    also = None
    if not args.no_also and name != "c2" and args.path == "fused":
        del r["prob"], r["evaluator"]
        prob = evaluator = None
        torch.cuda.empty_cache()
        a = measure(args, "c2", "weak", 200, 5, 50, dev, rank, world, local, windows, False)
        pk = peaks()
        also = {"workload": "c2: %s, weak-scaled: %d x 29,858 users" % (synth.CONFIGS["c2"]["name"], world), "scaling": "weak",
                "value": a["value"], "unit": UNIT, "ms_per_step": a["ms_per_step"], "steps": 200,
                "e2e_value": a["e2e"]["value"], "e2e_ms_per_step": a["e2e"]["ms_per_step"], "gpu_launches": a["launches"],
                "kernel": a.get("kernel"), "kernel_ms": a.get("kernel_ms"),
                "roofline_frac": (2.0 * a["U"] / world * a["I"] * a["d"] / (a["kernel_ms"] * 1e-3) / 1e12 /
                                  (pk["bf16"] if a.get("kernel") == "tcgen05_f16r" else pk["bf16"] / 2.0)) if a.get("kernel_ms") else None,
                "exact_rows": a.get("plan", {}).get("exact_rows"), "path": a["path"]}

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----------------------------------------------------------------
    pk = peaks()
    I, d = r["I"], r["d"]
    rows_rank0 = (U + world - 1) // world if name != "c5" else U
    items_rank0 = I if name != "c5" else (I + world - 1) // world
    roofline = None
    if args.path == "predict":
        # score-matrix-in: the kernel must read every score once, 4 I bytes per user
        sk = r["scores_kernel"]
        roofline = {"bound": "hbm", "kernel": "k_topk_scores + k_metrics + k_colsum_fold (skr_eval_scores)", "achieved": sk["gbs"], "peak": pk["hbm"],
                    "unit": "GB/s", "frac": sk["gbs"] / pk["hbm"], "traffic": None,
                    "algorithmic_bytes_per_launch": 4.0 * sk["rows"] * I, "kernel_ms": sk["ms"], "rows_per_launch": sk["rows"],
                    "note": "device-resident float32 [rows, I] block, CUDA events around skr_eval_scores, L2 flushed between launches"}
    elif "kernel_ms" in r:
        flops = 2.0 * rows_rank0 * items_rank0 * d  # rank 0's share of the step
        achieved = flops / (r["kernel_ms"] * 1e-3) / 1e12
        kern = r["kernel"]
        passes = {"tcgen05_3xtf32": 3}.get(kern, 1)
        if kern == "tcgen05_f16r":
            peak = pk["bf16"]
            peak_note = ("FP16 operands (tcgen05 kind::f16, FP32 accumulate): dense 16-bit peak = the %s cuBLAS bf16 burst peak (%.1f TFLOP/s) in "
                         "MEASURED_PEAKS.json; against the TF32 peak used for the tf32r kernel (half of it) the same throughput reads frac x 2" % (pk["src"], pk["bf16"]))
        elif kern.startswith("tcgen05"):
            peak = pk["bf16"] / 2.0
            peak_note = "TF32 dense = 1/2 of the %s cuBLAS bf16 burst peak (%.1f TFLOP/s) in MEASURED_PEAKS.json" % (pk["src"], pk["bf16"])
        else:
            peak = 2 * 128 * 148 * 1.965e9 / 1e12
            peak_note = "FP32 FMA pipe nominal (128 FMA/clk/SM x 148 SMs x 1965 MHz)"
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):  # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this config (ncu --set full)
            traffic = json.load(open(tp)).get("%s@%s@r2" % (kern, name))
        roofline = {"bound": "tensor", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_note": "per launch (one row chunk of <= 131,072 users) from profiles/traffic.json; null = no ncu capture of this kernel at this config in this round",
                    "kernel_ms": r["kernel_ms"], "kernel_launches_per_step": r["chunks_per_step"],
                    "kernel_share_of_step": r["kernel_ms"] / r["ms_per_step"], "prepass_kernel_ms": r["prepass_ms"], "plan": r["plan"],
                    "exact_rows_last_chunk": r["plan"].get("exact_rows"), "retried_rows_last_chunk": r["plan"].get("retried_rows"),
                    "algorithmic_flops_per_step_rank0": flops, "mma_passes": passes,
                    "tensor_pipe_utilisation_est": passes * achieved / peak, "peak_note": peak_note,
                    # the same kernel against the SUSTAINED cuBLAS figure (what a back-to-back GEMM holds at the power cap): the
                    # fair denominator for launches inside a long, power-capped step (c4, c5); `frac` keeps the burst peak
                    "frac_of_sustained_peak": (achieved / (pk["bf16_sustained"] * (peak / pk["bf16"]))) if kern.startswith("tcgen05") else None,
                    "rank0_rows": rows_rank0, "rank0_items": items_rank0}

    kern = r.get("kernel", "score blocks")
    dtype = {"tcgen05_3xtf32": "tf32x3 (fp32-grade)", "tcgen05_tf32r": "tf32 candidates + f32 re-scoring (fp32-exact)",
             "tcgen05_f16r": "scaled fp16 candidates (f32 accumulate) + f32 re-scoring (fp32-exact)",
             "tcgen05_1xtf32": "tf32", "simt_fp32": "f32"}.get(kern, "f32")
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": workload_config(name, scaling),
            "run": {"api": "RankingEvaluator.from_csr(..., shard_users=True%s).evaluate_device(model): the evaluate() of the product minus its final D2H copy, collective included"
                           % (", shard='items'" if name == "c5" else ""),
                    "parallelism": ("item-sharded x%d: per-rank top-K, all-gather of rank keys (overlapped), merge, metric-sum all-reduce" if name == "c5"
                                    else "user-sharded x%d, item table replicated, metric-sum all-reduce") % world,
                    "path": r["path"], "users_counted": r["users_counted"], "precision": args.precision,
                    "l2": "256 MB buffer written between steps, outside the event pairs",
                    "item_norms": args.norms, "clock_ramp_steps": r["n_ramp"], "prep_ms": r["prep_ms"], "data_generation_s": r["gen_s"]},
            "clocks": clocks, "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": roofline, "cpu_baseline": cpu,
            "metrics_at_top_k": {n_: float(v) for n_, v in zip(r["names"], r["means"])},
            "also": also}
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 10 (c4), 3 (c5), 200 (c1-c3)")
    ap.add_argument("--warmup", type=int, default=None, help="default 3 (c4, c5), 10 (c1-c3)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c3a", "c3b", "c4", "c5"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"], help="default: strong (c3a, c3b, c4, c5), weak (c1, c2)")
    ap.add_argument("--path", default="fused", choices=["fused", "predict"], help="predict: the model only offers the reference's predict()")
    ap.add_argument("--batch-size", type=int, default=256, help="user batch of the predict path (reference default 256)")
    ap.add_argument("--precision", default="auto", choices=["auto", "3xtf32", "fp32", "tf32r", "f16r"])
    ap.add_argument("--norms", default="iid", choices=["iid", "heavy"], help="heavy: heavy-tailed item norms, outliers, norm-correlated bias (c1-c3 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    args = ap.parse_args()
    big = args.config in LARGE
    if args.steps is None:
        args.steps = {"c4": 10, "c5": 3}.get(args.config, 200)
    if args.warmup is None:
        args.warmup = 3 if big else 10
    args.steps = max(1, args.steps)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
