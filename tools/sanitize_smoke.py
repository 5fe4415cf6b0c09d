"""Small shapes through every kernel path of the library, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitize_smoke.py

Covers: device-side CSR ingestion, fused tcgen05 (3xTF32, TF32 + re-scoring incl. the three-pass retry and the exact
per-row fallback), fused FP32 (dot and -L2 + bias), score blocks for top-K > 128, the score-matrix kernels (warp per
row, block per row), grouped column sums, per-shard lists + merge, row chunking.  Results are checked against each other
(tf32r == fp32 bit for bit), so a sanitizer-clean run is also a correct one.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

torch.cuda.set_device(0)
U, I, d, K = 300, 4000, 64, 20
dta = synth.make(users=U, items=I, d=d, nnz_train=6000, nnz_test=1500, seed=9, bias=True, norms="heavy", device="cuda")
ue, ie, b = (torch.from_numpy(dta[k]).cuda() for k in ("user_emb", "item_emb", "bias"))
metric = [1, 2, 3, 4, 5]
ctx = _native.Context(0)
ctx.set_train_csr(dta["train_indptr"], dta["train_indices"], I)
ctx.set_test_csr(dta["test_indptr"], dta["test_indices"], I)


def run(prec, k=K, **opts):
    for name, v in opts.items():
        ctx.set_option(name, v)
    idx = torch.empty((U, k), dtype=torch.int32, device="cuda")
    per = torch.empty((U, len(metric) * k), dtype=torch.float32, device="cuda")
    sums = torch.zeros(len(metric) * k, dtype=torch.float64, device="cuda")
    ctx.eval_fused(ue, ie, b, 0, metric, k, precision=prec, topk_idx=idx, per_user=per, sums=sums)
    torch.cuda.synchronize()
    for name in opts:
        ctx.set_option(name, {"retry_min": -1, "rank": 0, "sample_tiles": 0, "chunk_rows": 0, "score_fn": 0, "chunks": 0}[name])
    return idx.cpu().numpy(), per.cpu().numpy(), sums.cpu().numpy(), ctx.fused_stats(), ctx.last_fused_kernel


ref = run("fp32")
out = {"fp32": ref[4]}
for label, prec, opts in (("3xtf32", "3xtf32", {}), ("tf32r", "tf32r", {}), ("tf32r retry forced", "tf32r", {"retry_min": 1}),
                          ("tf32r useless thresholds -> exact rows", "tf32r", {"rank": 1, "sample_tiles": 2}),
                          ("tf32r chunked rows", "tf32r", {"chunk_rows": 128}), ("tf32r 3 item chunks", "tf32r", {"chunks": 3})):
    got = run(prec, **opts)
    same = np.array_equal(got[0], ref[0]) and (prec != "tf32r" or np.array_equal(got[1], ref[1]))
    out[label] = (got[4], got[3]["exact_rows"], got[3]["retried_rows"], bool(same))
    assert prec == "3xtf32" or same, label
out["neg_l2"] = run("fp32", score_fn=1)[4]
out["top-200 blocks"] = run("auto", k=200)[4]
# score-matrix kernels
S = (ue @ ie.T + b).contiguous()
sums = torch.zeros(len(metric) * K, dtype=torch.float64, device="cuda")
per = torch.empty((U, len(metric) * K), dtype=torch.float32, device="cuda")
ctx.eval_scores(S, 0, metric, K, per_user=per, sums=sums)
S2 = S[:, :3999].contiguous()  # odd pitch: misaligned rows
ctx.set_train_csr(None, None, 3999)
ctx.set_test_csr(dta["test_indptr"], np.minimum(dta["test_indices"], 3998), 3999)
ctx.eval_scores(S2, 0, metric, K, sums=sums)
idx = torch.empty((U, 300), dtype=torch.int32, device="cuda")
ctx.topk_scores(S2, 300, topk_idx=idx)  # block-per-row kernel (top-K > 128)
rows = torch.arange(0, U, 3, dtype=torch.int32, device="cuda")
gs = torch.zeros(len(metric) * K, dtype=torch.float64, device="cuda")
ctx.colsum_rows(per, rows, gs)
# per-shard lists + merge
ctx.set_train_csr(None, None, I)
ctx.set_test_csr(dta["test_indptr"], dta["test_indices"], I)
keys = torch.empty((2, U, K), dtype=torch.int64, device="cuda")
half = I // 2
c2 = _native.Context(0)
for sh, (lo, hi) in enumerate(((0, half), (half, I))):
    c2.set_train_csr(None, None, hi - lo)
    c2.topk_fused(ue, ie[lo:hi].contiguous(), b[lo:hi].contiguous(), 0, lo, K, keys[sh], precision="3xtf32")
ctx.eval_merged_topk(keys, 0, U, 0, metric, K, sums=sums)
torch.cuda.synchronize()
print("sanitize smoke ok:", out)
