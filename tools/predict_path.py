"""Where the time of the `predict` path goes (one B200): per 256-user batch of a c2-shaped model whose predict() is a GPU
GEMM followed by .cpu().numpy() (what an unmodified reference model does, LightGCN.py:214-216).

    python tools/predict_path.py [config=c2] [batch=256]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
I, d, K = cfg["items"], cfg["d"], max(cfg["top_k"])
g = torch.Generator(device=dev).manual_seed(1)
ue = torch.randn((4096, d), generator=g, device=dev) * 0.1
ie = torch.randn((I, d), generator=g, device=dev) * 0.1
ctx = _native.Context(0)
ctx.set_train_csr(None, None, I)
ctx.set_test_csr(np.arange(4097, dtype=np.int64), np.zeros(4096, np.int32), I)
ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
sums = torch.zeros(len(ids) * K, dtype=torch.float64, device=dev)


def wall(fn, n=10):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


users = torch.arange(B, device=dev)
out = {}
out["predict_gemm_ms"] = wall(lambda: ue[users] @ ie.T)
out["predict_total_ms (gemm + .cpu().numpy())"] = wall(lambda: (ue[users] @ ie.T).cpu().detach().numpy())
h = (ue[users] @ ie.T).cpu().numpy()
pin = torch.empty(h.size, dtype=torch.float32).pin_memory()
dbuf = torch.empty(h.size, dtype=torch.float32, device=dev)
out["stage_memcpy_ms (numpy -> pinned)"] = wall(lambda: pin.numpy().__setitem__(slice(None), h.ravel()))
out["h2d_pinned_ms"] = wall(lambda: dbuf.copy_(pin, non_blocking=True))
out["h2d_pageable_ms (torch.from_numpy(h).to(dev))"] = wall(lambda: torch.from_numpy(h).to(dev))
blk = dbuf.view(h.shape)
out["eval_scores_ms (mask + top-K + metrics, resident block)"] = wall(lambda: ctx.eval_scores(blk, 0, ids, K, sums=sums), 50)
rt = torch.cuda.cudart()


def registered():
    rt.cudaHostRegister(h.ctypes.data, h.nbytes, 0)
    dbuf.copy_(torch.from_numpy(h).view(-1), non_blocking=True)
    torch.cuda.synchronize()
    rt.cudaHostUnregister(h.ctypes.data)


out["h2d_register_copy_unregister_ms"] = wall(registered)
out["bytes_per_batch"] = int(h.nbytes)
# the HBM-bound kernel at a large resident block
for rows in (2048, 8192):
    if rows * I * 4 > (8 << 30):
        continue
    big = torch.randn((rows, I), generator=g, device=dev)
    ctx.set_test_csr(np.arange(rows + 1, dtype=np.int64), np.zeros(rows, np.int32), I)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ctx.eval_scores(big, 0, ids, K, sums=sums)
    for a, b in ev:
        flush.zero_()
        a.record()
        ctx.eval_scores(big, 0, ids, K, sums=sums)
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    out["eval_scores_%d_rows" % rows] = {"ms": ms, "GB/s": 4.0 * rows * I / (ms * 1e-3) / 1e9}
print(json.dumps({"config": name, "batch": B, "items": I, "K": K, **out}, indent=1))
