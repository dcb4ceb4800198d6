#!/usr/bin/env python
"""Per-layer timing of b2n_linear_wgrad at the training step's shapes (M = 327 040 samples) vs the library's dY^T X and vs the HBM bound."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from b2nerf._lib import lib
M = 327040
if os.environ.get('B2N_WG_LAYERS'):
    layers = [tuple(int(v) for v in t.split('x')) for t in os.environ['B2N_WG_LAYERS'].split(',')]
else:
  layers = [(64, 36), (32, 64), (16, 36), (1, 16), (32, 36), (1, 32), (64, 69), (64, 64), (65, 64), (64, 84), (3, 64)]
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
rows = []
st = torch.cuda.current_stream().cuda_stream
for n_out, n_in in layers:
    dy = torch.randn(M, n_out, device="cuda").half(); x = torch.randn(M, n_in, device="cuda").half()
    R = int(os.environ.get("B2N_WG_REPLICAS", "16"))
    dw = torch.zeros(R, n_out, n_in, device="cuda")
    us = t(lambda: lib().call("b2n_linear_wgrad_replicated", dy.data_ptr(), x.data_ptr(), M, n_out, n_in, dw.data_ptr(), R, 0, st))
    us_lib = t(lambda: dy.t() @ x, reps=5)
    byt = 2 * M * (n_out + n_in)
    rows.append({"out": n_out, "in": n_in, "us": round(us, 1), "library_us": round(us_lib, 1), "GBps": round(byt / us / 1e3, 1), "frac_hbm": round(byt / us / 1e3 / HBM, 3)})
    print(rows[-1], flush=True)
print(json.dumps({"M": M, "prefetch": os.environ.get("B2N_WGRAD_PREFETCH", "0"), "ctas_per_sm": os.environ.get("B2N_WGRAD_CTAS_PER_SM", "2"), "total_us": sum(r["us"] for r in rows), "library_total_us": sum(r["library_us"] for r in rows)}))
