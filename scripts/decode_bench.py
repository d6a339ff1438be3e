"""Decode timing on the GPU box: the round-1 workload (blank bias +1.0, every step has emitters) and the bench.py one
(bias +1.8), persistent kernel only, with the per-phase cycle counters."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rnnt_b200
import rnnt_b200.functional as RF

H = V = 1024
for bias, lo in ((1.0, 200), (1.8, 300)):
    torch.manual_seed(0)
    joint = rnnt_b200.JointNetwork(-1, -1, H, V)
    with torch.no_grad():
        joint.joint_ln.bias[V - 1] += bias
    model = rnnt_b200.RNNTModel(rnnt_b200.ConvPredictor(V, H, 512, 0.3), torch.nn.Identity(), joint).cuda().eval()
    feats = torch.randn(64, 400, H, device="cuda")
    lens = torch.randint(lo, 401, (64,)); lens[0] = 400
    model.greedy_decode_features(feats, lens, max_length=200)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        toks = model.greedy_decode_features(feats, lens, max_length=200)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    cyc = RF._last_decode_phase_cycles.tolist()
    steps = int(cyc[6])
    ref = model.greedy_decode_features(feats, lens, max_length=200, engine="graph")
    print(f"bias {bias}: {best*1e3:.2f} ms, ~{steps} steps, {best*1e6/steps:.1f} us/step, tokens {sum(len(x) for x in toks)}, "
          f"graph-engine identical: {ref == toks}")
    print("   phase cycles P1,P2,P3,conv1-table(once),P5,P6,steps,barriers:", cyc, " sum us @1.9GHz:", sum(cyc) / 1.9e3)
