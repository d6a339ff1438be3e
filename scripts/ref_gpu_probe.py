"""Context number: the reference's own call path (torch linear + tanh + torchaudio.functional.rnnt_loss, fp32 and
bf16-autocast GEMM) run ON THE B200 -- the 'existing sm_100 kernels' SURVEY 8d mentions -- next to the fused path, at
a batch the materialised logits allow.  Not part of bench.py: the contract's reference arm is the CPU path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torchaudio
from helpers import make_inputs
from rnnt_b200.functional import joint_rnnt_loss

def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for B in (8, 32):
    T, U, H, V = 400, 100, 1024, 1024
    inp = make_inputs(B, T, U, H, V)
    for k in ("enc", "pred", "W", "b"):
        inp[k].requires_grad_(True)
    cells = B * T * (U + 1)
    def ref(tf32=False):
        for k in ("enc", "pred", "W", "b"): inp[k].grad = None
        h = torch.tanh(inp["enc"].unsqueeze(2) + inp["pred"].unsqueeze(1))
        logits = torch.nn.functional.linear(h, inp["W"], inp["b"])
        loss = torchaudio.functional.rnnt_loss(logits, inp["targets"], inp["T_len"], inp["U_len"], blank=-1, clamp=-1, reduction="mean")
        loss.backward()
    def ours():
        for k in ("enc", "pred", "W", "b"): inp[k].grad = None
        joint_rnnt_loss(inp["enc"], inp["pred"], inp["W"], inp["b"], inp["targets"], inp["T_len"], inp["U_len"], validate=False).backward()
    torch.cuda.reset_peak_memory_stats()
    try:
        t_ref = timed(ref, 3)
        mem_ref = torch.cuda.max_memory_allocated() / 2**30
        torch.backends.cuda.matmul.allow_tf32 = True
        t_ref_tf32 = timed(ref, 3)
        torch.backends.cuda.matmul.allow_tf32 = False
    except torch.cuda.OutOfMemoryError:
        t_ref = t_ref_tf32 = float("nan"); mem_ref = float("nan")
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    t_ours = timed(ours, 10)
    mem_ours = torch.cuda.max_memory_allocated() / 2**30
    print(f"B={B}: reference path on the B200 fp32 {t_ref:.1f} ms ({cells/t_ref/1e3:.1f} M cells/s, peak {mem_ref:.1f} GiB), "
          f"TF32 matmul {t_ref_tf32:.1f} ms; fused path {t_ours:.2f} ms ({cells/t_ours/1e3:.1f} M cells/s, peak {mem_ours:.1f} GiB); "
          f"speed-up {t_ref/t_ours:.1f}x / {t_ref_tf32/t_ours:.1f}x", flush=True)
