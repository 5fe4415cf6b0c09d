"""skr_eval_scores on a resident [rows, I] block, a few launches (target of `ncu`): python tools/scores_only.py [config] [rows]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
cfg = synth.CONFIGS[name]
I, K = cfg["items"], max(cfg["top_k"])
ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
big = torch.randn((rows, I), generator=g, device=dev)
rng = np.random.default_rng(0)
ctx = _native.Context(0)
n_tr = 27
ctx.set_train_csr(np.arange(rows + 1, dtype=np.int64) * n_tr, rng.integers(0, I, size=rows * n_tr, dtype=np.int32), I)
ctx.set_test_csr(np.arange(rows + 1, dtype=np.int64) * 7, rng.integers(0, I, size=rows * 7, dtype=np.int32), I)
sums = torch.zeros(len(ids) * K, dtype=torch.float64, device=dev)
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    ctx.eval_scores(big, 0, ids, K, sums=sums)
torch.cuda.synchronize()
print("ok", name, rows, I, K)
