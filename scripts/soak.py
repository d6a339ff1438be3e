"""Soak test: many fused forward+backward steps over changing ragged batches; checks that repeated runs of the same
batch give bit-identical costs and finite gradients (catches rare synchronisation bugs that single runs miss)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_inputs
from rnnt_b200.functional import joint_rnnt_loss

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
shapes = [(32, 400, 100, 1024, 1024), (7, 123, 37, 256, 512), (16, 250, 60, 1024, 1024), (3, 33, 5, 64, 300)]
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < secs:
    B, T, U, H, V = shapes[n % len(shapes)]
    inp = make_inputs(B, T, U, H, V, ragged=True, seed=1000 + n)
    ref = None
    for rep in range(3):
        for k in ("enc", "pred", "W", "b"):
            inp[k].grad = None
            inp[k].requires_grad_(True)
        costs = joint_rnnt_loss(inp["enc"], inp["pred"], inp["W"], inp["b"], inp["targets"], inp["T_len"], inp["U_len"],
                                reduction="none", validate=False, save_hidden=(rep != 2))
        costs.sum().backward()
        c = costs.detach().clone()
        ok = bool(torch.isfinite(c).all()) and all(bool(torch.isfinite(inp[k].grad).all()) for k in ("enc", "pred", "W", "b"))
        if ref is None: ref = c
        elif not torch.equal(ref, c): ok = False
        if not ok:
            bad += 1
            print("MISMATCH at iteration", n, "rep", rep, (B, T, U, H, V), flush=True)
    n += 1
torch.cuda.synchronize()
print(f"soak: {n} batches x 3 runs in {time.time() - t0:.1f} s, failures: {bad}", flush=True)
sys.exit(1 if bad else 0)
