"""cuBLAS TF32 throughput on this box (SURVEY 8d asks for it next to the derived peak = bf16 / 2):
8192^3 torch.matmul with allow_tf32, CUDA events, best and median of 20 after warm-up; bf16 for comparison."""
import json

import numpy as np
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
out = {}
for name, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    a = torch.randn((n, n), device="cuda", dtype=dt)
    b = torch.randn((n, n), device="cuda", dtype=dt)
    for _ in range(10):
        a @ b
    torch.cuda.synchronize()
    ms = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    fl = 2.0 * n ** 3
    out[name] = {"best_tflops": fl / (min(ms) * 1e-3) / 1e12, "median_tflops": fl / (float(np.median(ms)) * 1e-3) / 1e12}
print(json.dumps({"cublas_8192_cubed": out}))
