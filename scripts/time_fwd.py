"""Time the forward / backward C-ABI calls at the bench shape (diagnostics)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_inputs
import rnnt_b200
from rnnt_b200.functional import joint_rnnt_loss

B, T, U, H, V = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (32, 400, 100, 1024, 1024))]
inp = make_inputs(B, T, U, H, V)
for k in ("enc", "pred", "W", "b"):
    inp[k].requires_grad_(True)
def fwd():
    return joint_rnnt_loss(inp["enc"], inp["pred"], inp["W"], inp["b"], inp["targets"], inp["T_len"], inp["U_len"], validate=False)
for mode in ("fwd", "fwdbwd"):
    for _ in range(3):
        l = fwd()
        if mode == "fwdbwd": l.backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        l = fwd()
        if mode == "fwdbwd": l.backward()
    e1.record(); torch.cuda.synchronize()
    print(f"DBG={os.environ.get('RNNT_B200_DBG','0')} {mode}: {e0.elapsed_time(e1)/n:.3f} ms  loss {float(l.detach()):.4f}", flush=True)
