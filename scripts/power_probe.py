"""Run the forward (or full step) in a loop for a few seconds while sampling nvidia-smi clocks/power (diagnostics)."""
import os, sys, time, subprocess, threading, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import make_inputs
from rnnt_b200.functional import joint_rnnt_loss
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
inp = make_inputs(32, 400, 100, 1024, 1024)
for k in ("enc", "pred", "W", "b"):
    inp[k].requires_grad_(True)
def step():
    l = joint_rnnt_loss(inp["enc"], inp["pred"], inp["W"], inp["b"], inp["targets"], inp["T_len"], inp["U_len"], validate=False)
    if mode == "fwdbwd":
        l.backward()
# warm up (and fill the activation buffer) with the full kernels, then switch the diagnostic stub bits on
dbg = os.environ.pop("RNNT_B200_DBG", "0")
for _ in range(3): step()
torch.cuda.synchronize()
os.environ["RNNT_B200_DBG"] = dbg
lines = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
th = threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True); th.start()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    for _ in range(10): step()
    n += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
proc.terminate()
ms = e0.elapsed_time(e1) / n
vals = [l.strip().split(",") for l in lines if l.count(",") >= 2]
clk = [float(v[0]) for v in vals]; pw = [float(v[1]) for v in vals]
half = len(clk) // 2
print(f"DBG={os.environ.get('RNNT_B200_DBG','0')} {mode}: {ms:.3f} ms/iter over {n} iters; clocks median {statistics.median(clk[half:]):.0f} MHz (min {min(clk):.0f}), power median {statistics.median(pw[half:]):.0f} W max {max(pw):.0f} W, power_cap active in {sum('Active' in v[2] for v in vals)}/{len(vals)} samples", flush=True)
