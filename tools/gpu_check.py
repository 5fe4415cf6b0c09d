"""Staged GPU diagnostics (development aid, not part of the product): each stage runs in its own
subprocess under a timeout so that a faulting kernel cannot take the others down.
    python tools/gpu_check.py [stage ...]      -> gpurun_out/check_<stage>.log
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
STAGES = ["scores", "simt", "tc_probe", "tc_small", "tc_c2"]


def _setup():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import oracle
    from skrec_b200 import _native
    return np, torch, oracle, _native


def _csr(np, g, B, N, mx, mn=0):
    sizes = g.integers(mn, mx + 1, size=B)
    ptr = np.zeros(B + 1, np.int64)
    np.cumsum(sizes, out=ptr[1:])
    rows = [g.choice(N, size=int(n), replace=False) for n in sizes]
    return ptr, (np.concatenate(rows) if ptr[-1] else np.zeros(0)).astype(np.int32)


def _fused(np, torch, ctx, ue, ie, b, tr, te, metric, K, prec):
    U, I = ue.shape[0], ie.shape[0]
    ctx.set_train_csr(tr[0], tr[1], I) if tr is not None else ctx.set_train_csr(None, None, I)
    ctx.set_test_csr(te[0], te[1], I)
    MK = len(metric) * K
    idx = torch.empty((U, K), dtype=torch.int32, device="cuda")
    val = torch.empty((U, K), dtype=torch.float32, device="cuda")
    per = torch.empty((U, MK), dtype=torch.float32, device="cuda")
    sums = torch.zeros(MK, dtype=torch.float64, device="cuda")
    ctx.eval_fused(torch.from_numpy(ue).cuda(), torch.from_numpy(ie).cuda(), None if b is None else torch.from_numpy(b).cuda(),
                   0, metric, K, precision=prec, topk_idx=idx, topk_val=val, per_user=per, sums=sums)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), val.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy(), ctx.fused_kernel_ms(0)


def _report(np, oracle, name, out, ue, ie, b, tr, te, metric, K):
    idx, val, per, sums, ms = out
    S = oracle.scores(ue, ie, b)
    if tr is not None:
        oracle.mask_rows(S, tr[0], tr[1])
    eper, etop = oracle.eval_scores(S, te[0], te[1], metric, K, return_topk=True)
    gs = np.take_along_axis(S, np.clip(idx, 0, S.shape[1] - 1).astype(np.int64), 1)
    es = np.take_along_axis(S, etop.astype(np.int64), 1)
    fin = np.isfinite(gs) & (idx >= 0)
    diff = idx != etop
    U = ue.shape[0]
    print("%s: kernel %.3f ms | idx mismatch %.4f%% | max|val - S[idx]| %.3e | max near-tie gap %.3e | max mean-metric diff %.3e | bad idx %d"
          % (name, ms, 100 * diff.mean(), float(np.max(np.abs(val[fin] - gs[fin]))) if fin.any() else -1,
             float(np.max(np.abs(gs[diff] - es[diff]))) if diff.any() else 0.0,
             float(np.max(np.abs(sums / U - oracle.sums_f64(eper) / U))), int((idx < 0).sum())), flush=True)


def stage_scores():
    np, torch, oracle, _native = _setup()
    ctx = _native.Context(0)
    g = np.random.default_rng(0)
    for (B, N, K) in [(64, 257, 10), (33, 4099, 50), (200, 40981, 50)]:
        s = np.stack([(g.permutation(N).astype(np.float32) - N / 2) / np.float32(N) for _ in range(B)])
        tr, te = _csr(np, g, B, N, 50), _csr(np, g, B, N, 20)
        ctx.set_train_csr(tr[0], tr[1], N)
        ctx.set_test_csr(te[0], te[1], N)
        per = torch.empty((B, 5 * K), dtype=torch.float32, device="cuda")
        idx = torch.empty((B, K), dtype=torch.int32, device="cuda")
        ctx.eval_scores(torch.from_numpy(s).cuda(), 0, [1, 2, 3, 4, 5], K, topk_idx=idx, per_user=per)
        torch.cuda.synchronize()
        m = s.copy()
        oracle.mask_rows(m, tr[0], tr[1])
        eper, etop = oracle.eval_scores(m, te[0], te[1], [1, 2, 3, 4, 5], K, return_topk=True)
        print("scores B=%d N=%d K=%d: idx equal %s, per_user equal %s" % (B, N, K, np.array_equal(idx.cpu().numpy(), etop),
                                                                         np.array_equal(per.cpu().numpy(), eper)), flush=True)


def _small_cases(np):
    g = np.random.default_rng(1)
    for (U, I, d, bias, K) in [(128, 128, 32, False, 128), (300, 1000, 64, True, 10), (257, 4097, 128, True, 100), (1000, 5000, 32, False, 20)]:
        ue = (g.standard_normal((U, d)) * 0.1).astype(np.float32)
        ie = (g.standard_normal((I, d)) * 0.1).astype(np.float32)
        b = (g.standard_normal(I) * 0.01).astype(np.float32) if bias else None
        tr = None if K == 128 else _csr(np, g, U, I, 40)
        te = _csr(np, g, U, I, 20, 1)
        yield U, I, d, K, ue, ie, b, tr, te


def stage_simt():
    np, torch, oracle, _native = _setup()
    ctx = _native.Context(0)
    for U, I, d, K, ue, ie, b, tr, te in _small_cases(np):
        out = _fused(np, torch, ctx, ue, ie, b, tr, te, [1, 2, 4], K, "fp32")
        _report(np, oracle, "simt U=%d I=%d d=%d K=%d" % (U, I, d, K), out, ue, ie, b, tr, te, [1, 2, 4], K)


def stage_tc_probe():
    """One 128x128 tile, K=128: the sorted list is the whole score row, so the raw accumulator is visible."""
    np, torch, oracle, _native = _setup()
    ctx = _native.Context(0)
    g = np.random.default_rng(2)
    for d in (32, 64, 128):
        for prec in ("1xtf32", "3xtf32"):
            U = I = 128
            ue = (g.standard_normal((U, d))).astype(np.float32)
            ie = (g.standard_normal((I, d))).astype(np.float32)
            te = _csr(np, g, U, I, 5, 1)
            idx, val, per, sums, ms = _fused(np, torch, ctx, ue, ie, None, None, te, [1], 128, prec)
            got = np.full((U, I), np.nan, np.float32)
            ok = idx >= 0
            rows = np.repeat(np.arange(U), 128).reshape(U, 128)
            got[rows[ok], idx[ok]] = val[ok]
            exact = (ue.astype(np.float64) @ ie.astype(np.float64).T)
            err = np.abs(got - exact)
            print("probe d=%d %s: %.3f ms, filled %d/%d, max err %.3e, mean err %.3e, rows with err>1e-2: %d, cols with err>1e-2: %d"
                  % (d, prec, ms, int(np.isfinite(got).sum()), U * I, float(np.nanmax(err)), float(np.nanmean(err)),
                     int((np.nanmax(err, 1) > 1e-2).sum()), int((np.nanmax(err, 0) > 1e-2).sum())), flush=True)
            if np.nanmax(err) > 1e-2:
                np.save(os.path.join(OUT, "probe_got_d%d_%s.npy" % (d, prec)), got)
                np.save(os.path.join(OUT, "probe_exact_d%d_%s.npy" % (d, prec)), exact.astype(np.float32))


def stage_tc_small():
    np, torch, oracle, _native = _setup()
    ctx = _native.Context(0)
    for U, I, d, K, ue, ie, b, tr, te in _small_cases(np):
        out = _fused(np, torch, ctx, ue, ie, b, tr, te, [1, 2, 4], K, "3xtf32")
        _report(np, oracle, "tc3 U=%d I=%d d=%d K=%d" % (U, I, d, K), out, ue, ie, b, tr, te, [1, 2, 4], K)
        print("      prepass %.3f ms, plan %s" % (ctx.fused_prepass_ms(0), ctx.fused_stats()), flush=True)


def stage_tc_c2():
    np, torch, oracle, _native = _setup()
    from skrec_b200 import synth
    ctx = _native.Context(0)
    d = synth.make_config("c2", device="cuda")
    tr = (d["train_indptr"], d["train_indices"])
    te = (d["test_indptr"], d["test_indices"])
    res = {}
    for prec in ("3xtf32", "fp32", "1xtf32", "tf32r"):
        for chunks in ((0, 2, 4, 5, 8) if prec == "3xtf32" else (0,)):
            ctx.set_option("chunks", chunks)
            out = _fused(np, torch, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, prec)
            out = _fused(np, torch, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, prec)
            res[(prec, chunks)] = out
            print("c2 %s chunks=%d: kernel %.3f ms (+ prepass %.3f) -> %.1f TFLOP/s algorithmic; NDCG@50 %.6f; %s" % (
                prec, chunks, out[4], ctx.fused_prepass_ms(0), 2.0 * d["users"] * d["items"] * 64 / out[4] / 1e9,
                out[3][149] / d["users"], ctx.fused_stats() if prec != "fp32" else ""), flush=True)
    ctx.set_option("chunks", 0)
    for st, rk in ((16, 0), (64, 0), (0, 16), (0, 32)):
        ctx.set_option("sample_tiles", st)
        ctx.set_option("rank", rk)
        out = _fused(np, torch, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, "3xtf32")
        out = _fused(np, torch, ctx, d["user_emb"], d["item_emb"], None, tr, te, [1, 2, 4], 50, "3xtf32")
        print("c2 3xtf32 sample_tiles=%d rank=%d: kernel %.3f ms (+ prepass %.3f); NDCG@50 %.6f; %s" % (
            st, rk, out[4], ctx.fused_prepass_ms(0), out[3][149] / d["users"], ctx.fused_stats()), flush=True)
    ctx.set_option("sample_tiles", 0)
    ctx.set_option("rank", 0)
    a, b = res[("tf32r", 0)], res[("fp32", 0)]
    print("c2 tf32r vs simt fp32: idx equal %s, val equal %s, per-user equal %s, max sum diff %.3e" % (
        np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1]), np.array_equal(a[2], b[2]), float(np.max(np.abs(a[3] - b[3])))), flush=True)
    a, b = res[("3xtf32", 0)], res[("fp32", 0)]
    diff = a[0] != b[0]
    print("c2 tc3 vs simt: idx mismatch %.4f%%, max val gap at mismatches %.3e, max mean-metric diff %.3e" % (
        100 * diff.mean(), float(np.max(np.abs(a[1][diff] - b[1][diff]))) if diff.any() else 0.0,
        float(np.max(np.abs(a[3] - b[3])) / d["users"])), flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        globals()["stage_" + sys.argv[2]]()
        return 0
    stages = sys.argv[1:] or STAGES
    rc_all = 0
    for st in stages:
        log = os.path.join(OUT, "check_%s.log" % st)
        with open(log, "w") as f:
            try:
                rc = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], stdout=f, stderr=subprocess.STDOUT,
                                    timeout=240).returncode
            except subprocess.TimeoutExpired:
                rc = -999
                f.write("\nTIMEOUT\n")
        print("== stage %s rc=%d" % (st, rc))
        print(open(log).read()[-3000:])
        rc_all |= (rc != 0)
    return rc_all


if __name__ == "__main__":
    sys.exit(main())
