"""Bring-up ladder for the GPU box: runs each stage in its own process with a timeout and prints diagnostics.

    python scripts/gpu_ladder.py [stage ...]     (stages: dense decode fwd bwd all)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage_dense():
    import torch, torchaudio
    import rnnt_b200
    from helpers import make_inputs, rel_err
    for (B, T, U, V, ragged) in [(3, 13, 6, 37, True), (4, 50, 20, 1024, True), (2, 400, 100, 64, False)]:
        inp = make_inputs(B, T, U, 16, V, ragged=ragged)
        logits = torch.randn(B, T, U + 1, V, device="cuda").requires_grad_(True)
        l2 = logits.detach().clone().requires_grad_(True)
        mine = rnnt_b200.rnnt_loss(logits, inp["targets"], inp["T_len"], inp["U_len"], reduction="none")
        ref = torchaudio.functional.rnnt_loss(l2, inp["targets"], inp["T_len"], inp["U_len"], reduction="none")
        mine.sum().backward(); ref.sum().backward()
        print("dense", (B, T, U, V), "cost rel", rel_err(mine, ref), "grad", rel_err(logits.grad, l2.grad), flush=True)


def stage_decode():
    import torch
    import rnnt_b200
    from helpers import rel_err
    torch.manual_seed(0)
    for (N, H, V) in [(5, 64, 48), (64, 1024, 1024)]:
        a, p = torch.randn(N, H, device="cuda"), torch.randn(N, H, device="cuda")
        W, b = torch.randn(V, H, device="cuda") / H ** 0.5, torch.randn(V, device="cuda") * 0.1
        tok, mg = rnnt_b200.joint_argmax(a, p, W, b, return_margin=True)
        ref = (torch.tanh(a.double() + p.double()) @ W.double().T + b.double())
        top = ref.topk(2, -1).values
        print("decode", (N, H, V), "mismatch", int((tok.long() != ref.argmax(-1)).sum()),
              "margin err", float((mg.double() - (top[:, 0] - top[:, 1])).abs().max()), flush=True)


def _fwd_bwd(B, T, U, H, V, ragged, ring_tiles=None, check_rings=True):
    import torch
    from helpers import make_inputs, fused_raw, torch_reference, rel_err, ring_views, tile_rows
    inp = make_inputs(B, T, U, H, V, ragged=ragged)
    ref = torch_reference(inp)
    refq = torch_reference(inp, emulate_bf16=True)
    out = fused_raw(inp, ring_tiles=ring_tiles, flags=int(os.environ.get('BWD_FLAGS', '0')))
    tag = f"B{B} T{T} U{U} H{H} V{V} ragged={ragged} ring={out['ring_tiles']}"
    print(tag, "status", out["status"])
    print("  costs mine", out["costs"][:4].tolist(), "ref", ref["costs"][:4].tolist())
    print("  costs rel vs fp32", rel_err(out["costs"], ref["costs"]), "vs bf16-emulated", rel_err(out["costs"], refq["costs"]))
    # lse / lp check on valid cells against fp32 logits
    lse_ref = torch.logsumexp(refq["logits"], -1)
    mask = torch.zeros_like(lse_ref, dtype=torch.bool)
    for b in range(B):
        mask[b, : int(inp["T_len"][b]), : int(inp["U_len"][b]) + 1] = True
    print("  lse max abs err (valid)", float((out["lse"] - lse_ref)[mask].abs().max()))
    lpB_ref = refq["logits"][..., -1] - lse_ref
    print("  lpB max abs err (valid)", float((out["lp"][..., 0] - lpB_ref)[mask].abs().max()))
    for k in ("d_enc", "d_pred", "dW", "db"):
        print(f"  {k}: vs fp32 {rel_err(out[k], ref[k])}  vs bf16-emulated {rel_err(out[k], refq[k])}", flush=True)
    if check_rings and out["ring_tiles"] * 128 <= 200000:
        g_ring, h_ring, S = ring_views(out)
        print('  gradient scale S =', S)
        rows = tile_rows(inp["T_len"], inp["U_len"], out)
        print("  active half-tiles", out["active_halves"], "of", out["total_halves"])
        n = min(len(rows), g_ring.shape[0])
        bi = torch.tensor([r[0] for r in rows[:n]], device="cuda")
        ti = torch.tensor([min(r[1], T - 1) for r in rows[:n]], device="cuda")
        ui = torch.tensor([min(r[2], U) for r in rows[:n]], device="cuda")
        vd = torch.tensor([r[3] for r in rows[:n]], device="cuda")
        h_ref = torch.tanh(inp["enc"][bi, ti] + inp["pred"][bi, ui])
        herr = (h_ring[:n, :H].float() - h_ref)[vd]
        print("  h_ring max abs err (valid rows)", float(herr.abs().max()))
        # expected g from the bf16-emulated logits and our own alpha/beta
        import torchaudio  # noqa
        lg = refq["logits"][bi, ti, ui]
        al, be = out["alpha"], out["beta"]
        logZ = be[:, 0, 0][bi]
        p = torch.softmax(lg, -1)
        gam = torch.exp(al[bi, ti, ui] + be[bi, ti, ui] - logZ)
        g_ref = p * gam[:, None]
        Tb = inp["T_len"][bi].long(); Ub = inp["U_len"][bi].long()
        lpB = out["lp"][bi, ti, ui, 0]; lpE = out["lp"][bi, ti, ui, 1]
        c = al[bi, ti, ui] - logZ
        be_dn = be[bi, torch.clamp(ti + 1, max=T - 1), ui]
        be_rt = be[bi, ti, torch.clamp(ui + 1, max=U)]
        eB = torch.where(ti < Tb - 1, torch.exp(c + lpB + be_dn), torch.where(ui == Ub, torch.exp(c + lpB), torch.zeros_like(c)))
        eE = torch.where(ui < Ub, torch.exp(c + lpE + be_rt), torch.zeros_like(c))
        g_ref[:, V - 1] -= eB
        tg = inp["targets"].long()[bi, torch.clamp(ui, max=U - 1)]
        g_ref[torch.arange(n, device="cuda"), tg] -= eE
        g_ref = torch.where(vd[:, None], g_ref, torch.zeros_like(g_ref))
        gerr = g_ring[:n, :V].float() / S - g_ref
        print("  g_ring rel err", float(gerr.norm() / g_ref.norm()), "max abs", float(gerr.abs().max()),
              "invalid-row max", float(g_ring[:n][~vd].float().abs().max()) if (~vd).any() else 0.0)
        # downstream expectations computed FROM the rings (isolates dh / dW kernels)
        gf, hf = g_ring[:n].float() / S, h_ring[:n].float()
        Wb = torch.zeros(out["Vp"], out["Hp"], device="cuda"); Wb[:V, :H] = inp["W"].half().float()
        dW_exp = (gf.T @ hf)[:V, :H]
        print("  dW vs rings-expected", rel_err(out["dW"], dW_exp), " db vs rings", rel_err(out["db"], gf.sum(0)[:V]))
        hx = torch.zeros_like(hf); hx[:, :H] = h_ref
        dz = (gf @ Wb) * (1 - hx * hx)
        d_enc_exp = torch.zeros_like(out["d_enc"]); d_pred_exp = torch.zeros_like(out["d_pred"])
        d_enc_exp.index_put_((bi[vd], ti[vd]), dz[vd][:, :H], accumulate=True)
        d_pred_exp.index_put_((bi[vd], ui[vd]), dz[vd][:, :H], accumulate=True)
        print("  d_enc vs rings-expected", rel_err(out["d_enc"], d_enc_exp), " d_pred", rel_err(out["d_pred"], d_pred_exp), flush=True)


def stage_fwd():
    _fwd_bwd(2, 20, 9, 64, 256, True)


def stage_bwd():
    _fwd_bwd(3, 37, 13, 128, 512, True)
    _fwd_bwd(2, 33, 17, 1024, 1024, False)
    _fwd_bwd(2, 40, 20, 256, 1024, True, ring_tiles=7)   # multi-chunk backward


def stage_big():
    _fwd_bwd(4, 200, 40, 1024, 1024, False, check_rings=False)


STAGES = dict(dense=stage_dense, decode=stage_decode, fwd=stage_fwd, bwd=stage_bwd, big=stage_big)

if __name__ == "__main__":
    args = sys.argv[1:] or ["all"]
    if args[0] == "--run":
        STAGES[args[1]]()
        sys.exit(0)
    names = list(STAGES) if args == ["all"] else args
    for name in names:
        print(f"===== stage {name}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], timeout=240,
                               capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"[stage {name} exit {r.returncode}]\n" + r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            print(f"[stage {name} TIMEOUT]", (e.stdout or b"")[-2000:], (e.stderr or b"")[-2000:])
