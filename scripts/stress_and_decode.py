"""BASELINE configs[3] (long-utterance, large-vocab stress) and configs[4] (batched greedy decode) on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import rnnt_b200
import rnnt_b200.functional as RF
from helpers import make_inputs

def stress():
    B, T, U, H, V = 8, 1500, 300, 1024, 4096
    inp = make_inputs(B, T, U, H, V, ragged=True, seed=5)
    for k in ("enc", "pred", "W", "b"):
        inp[k].requires_grad_(True)
    torch.cuda.reset_peak_memory_stats()
    RF.COLLECT_BACKWARD_STATS = True
    def step():
        for k in ("enc", "pred", "W", "b"):
            inp[k].grad = None
        costs, res = rnnt_b200.joint_rnnt_loss(inp["enc"], inp["pred"], inp["W"], inp["b"], inp["targets"], inp["T_len"],
                                               inp["U_len"], reduction="none", validate=False, return_residuals=True)
        costs.sum().backward()
        return costs, res
    costs, res = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        costs, res = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    cells = int(((inp["T_len"].long()) * (inp["U_len"].long() + 1)).sum())
    act, tot = RF.last_backward_stats()
    tail = []
    for b in range(B):
        Tb, Ub = int(inp["T_len"][b]), int(inp["U_len"][b])
        tail.append(-(res["alpha"][b, Tb - 1, Ub] + res["lp"][b, Tb - 1, Ub, 0]))
    tail = torch.stack(tail)
    print(f"stress B{B} T{T} U{U} V{V}: {ms:.2f} ms/step, valid cells {cells} -> {cells/ms/1e3:.1f} M cells/s; "
          f"active tiles {act}/{tot}; peak mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB "
          f"(reference would need {B*T*(U+1)*V*4*2/2**30:.0f} GiB for logits+grads)")
    print("  costs", [round(float(c), 2) for c in costs], "alpha/beta consistency max rel",
          float(((tail - costs).abs() / costs.abs()).max()))
    print("  grads finite:", all(torch.isfinite(inp[k].grad).all().item() for k in ("enc", "pred", "W", "b")),
          " sum(db)/sum|db| =", float(inp["b"].grad.sum() / inp["b"].grad.abs().sum()))
    # sub-problem check: utterance with the shortest length alone, against torchaudio on the GPU (fp32)
    from helpers import torch_reference, rel_err
    small = make_inputs(1, 120, 40, H, V, seed=9)
    ref = torch_reference(small)
    e = small["enc"].clone().requires_grad_(True)
    c = rnnt_b200.joint_rnnt_loss(e, small["pred"], small["W"], small["b"], small["targets"], small["T_len"], small["U_len"], reduction="none")
    c.sum().backward()
    print("  V=4096 sub-problem: cost rel", rel_err(c.detach(), ref["costs"])[0], "d_enc rel", rel_err(e.grad, ref["d_enc"])[0])

def decode():
    torch.manual_seed(0)
    B, T, H, V, E = 64, 400, 1024, 1024, 512
    joint = rnnt_b200.JointNetwork(-1, -1, H, V)
    pred = rnnt_b200.ConvPredictor(V, H, E, 0.3)
    with torch.no_grad():
        joint.joint_ln.bias[V - 1] += 1.0      # random init emits on almost every frame; bias towards blank a bit
    model = rnnt_b200.RNNTModel(pred, torch.nn.Identity(), joint).cuda().eval()
    feats = torch.randn(B, T, H, device="cuda")
    lens = torch.randint(T // 2, T + 1, (B,))
    lens[0] = T
    out = model.greedy_decode_features(feats, lens, max_length=200)
    out, margins = model.greedy_decode_features(feats, lens, max_length=200, return_margins=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out2 = model.greedy_decode_features(feats, lens, max_length=200)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert out2 == out
    cyc = RF._last_decode_phase_cycles.tolist()
    print('  decode kernel phase cycles (block 0): P1 %d P2 %d P3 %d P4 %d P5 %d P6 %d barriers %d' % (cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], cyc[5], cyc[7]), ' total us @1.9GHz: %.0f' % (sum(cyc)/1900))
    steps = max(len(m) for m in margins)
    t1 = time.perf_counter()
    out_h = model._greedy_decode_features_hostloop(feats, lens, max_length=200)
    torch.cuda.synchronize()
    print(f"  host-driven loop: {(time.perf_counter()-t1)*1e3:.1f} ms, identical tokens: {out_h == out}")
    t1 = time.perf_counter()
    out_e = model.greedy_decode_features(feats, lens, max_length=200, use_cuda_graph=False, engine="graph")
    torch.cuda.synchronize()
    print(f"  eager device loop: {(time.perf_counter()-t1)*1e3:.1f} ms, identical tokens: {out_e == out}")
    model.greedy_decode_features(feats, lens, max_length=200, engine="graph")
    t1 = time.perf_counter()
    out_g = model.greedy_decode_features(feats, lens, max_length=200, engine="graph")
    torch.cuda.synchronize()
    print(f"  CUDA-graph device loop: {(time.perf_counter()-t1)*1e3:.1f} ms, identical tokens: {out_g == out}")
    print(f"decode (persistent kernel) B{B} T{T}: {dt*1e3:.1f} ms total, {steps} batched steps ({dt/steps*1e6:.0f} us/step), "
          f"tokens/utt mean {sum(len(o) for o in out)/B:.1f}, frames/s {int(lens.sum())/dt:.0f}, min top-2 margin {min(min(m) for m in margins):.2e}")
    # reference algorithm, utterance by utterance (rnnt/model.py:90-128 restated with torch ops on the GPU, fp32)
    mism = 0
    for b in range(0, B, 8):
        toks = [V - 1]; t = 0; per = 0
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            pf = model.predictor(torch.tensor([toks], device="cuda"))
            while t < int(lens[b]) and len(toks) < 200:
                lg = model.joint.single_forward(feats[b:b+1, t, :], pf[:, -1, :])
                tok = int(lg.argmax(-1))
                if tok == V - 1 or per >= 10:
                    t += 1; per = 0
                else:
                    toks.append(tok); pf = model.predictor(torch.tensor([toks], device="cuda")); per += 1
        mism += int(toks[1:] != out[b])
    print("  utterances checked against the per-utterance reference loop:", B // 8, "mismatching:", mism)

if __name__ == "__main__":
    which = sys.argv[1:] or ["stress", "decode"]
    if "stress" in which: stress()
    if "decode" in which: decode()
