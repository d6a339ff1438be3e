"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py          # needs /root/reference (absent on the GPU box)

Imports rnnt.joint.JointNetwork / rnnt.model.RNNTModel / rnnt.predictor.ConvPredictor from
/root/reference and torchaudio.functional.rnnt_loss (the call at rnnt/model.py:35-41), runs them in
fp32 on CPU with fixed seeds, and writes small .npz fixtures next to this script.  The fixtures are
committed; tests never read /root/reference.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import torchaudio  # noqa: E402
from rnnt.joint import JointNetwork  # noqa: E402
from rnnt.model import RNNTModel  # noqa: E402
from rnnt.predictor import ConvPredictor  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def loss_case(name, B, T, U, H, V, T_len, U_len, seed):
    torch.manual_seed(seed)
    joint = JointNetwork(-1, -1, H, V)
    enc = torch.randn(B, H, T).permute(0, 2, 1).contiguous().requires_grad_(True)   # model.py:28 layout, made dense
    pred = torch.randn(B, U + 1, H, requires_grad=True)
    targets = torch.randint(0, V - 1, (B, U), dtype=torch.int32)
    T_len = torch.tensor(T_len, dtype=torch.int32)
    U_len = torch.tensor(U_len, dtype=torch.int32)
    logits = joint(enc, pred)                                                     # joint.py:25-39
    logits.retain_grad()
    costs = torchaudio.functional.rnnt_loss(logits=logits, targets=targets, logit_lengths=T_len,
                                            target_lengths=U_len, blank=-1, clamp=-1, reduction="none")
    costs.sum().backward()
    mean = torchaudio.functional.rnnt_loss(logits=logits.detach(), targets=targets, logit_lengths=T_len,
                                           target_lengths=U_len, blank=-1, clamp=-1, reduction="mean")
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        enc=enc.detach().numpy(), pred=pred.detach().numpy(),
        W=joint.joint_ln.weight.detach().numpy(), b=joint.joint_ln.bias.detach().numpy(),
        targets=targets.numpy(), T_len=T_len.numpy(), U_len=U_len.numpy(),
        costs=costs.detach().numpy(), loss_mean=mean.numpy(),
        d_enc=enc.grad.numpy(), d_pred=pred.grad.numpy(),
        dW=joint.joint_ln.weight.grad.numpy(), db=joint.joint_ln.bias.grad.numpy(),
        logits=logits.detach().numpy().astype(np.float32) if logits.numel() < 200000 else np.zeros(0, np.float32),
        dlogits=logits.grad.numpy() if logits.numel() < 200000 else np.zeros(0, np.float32),
    )
    print(name, "costs", costs.detach().numpy())


class FixedEncoder(torch.nn.Module):
    """Stands in for AudioEncoder: returns the given (N,C,L) features (model.py:93-95 consumes them)."""
    def __init__(self, feats):
        super().__init__()
        self.feats = feats
        self.p = torch.nn.Parameter(torch.zeros(1))

    def forward(self, mel):
        return self.feats

    def calc_output_lens(self, lens):
        return lens


def decode_case(name, n_utt, T, H, V, E, seed, max_length, scale, blank_bias):
    torch.manual_seed(seed)
    joint = JointNetwork(-1, -1, H, V)
    predictor = ConvPredictor(V, H, E, 0.3).eval()
    # make non-blank emissions likely enough that the loop exercises emits and the 10-per-frame cap
    with torch.no_grad():
        joint.joint_ln.weight.mul_(scale)
        joint.joint_ln.bias[V - 1] += blank_bias
    feats = torch.randn(n_utt, T, H)
    T_len = torch.randint(T // 2, T + 1, (n_utt,))
    T_len[0] = T
    tokens = []
    for i in range(n_utt):
        f = feats[i, : int(T_len[i])].t().unsqueeze(0).contiguous()                # (1,C,L)
        model = RNNTModel(predictor, FixedEncoder(f), joint).eval()
        toks = model.greedy_decode(torch.zeros(1, 1, 1), torch.tensor([int(T_len[i])]), max_length=max_length)
        tokens.append(np.asarray(toks, np.int32))
        print(name, i, int(T_len[i]), len(toks), toks[:12])
    sd = {("pred." + k): v.numpy() for k, v in predictor.state_dict().items()}
    flat = np.concatenate(tokens) if tokens else np.zeros(0, np.int32)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), feats=feats.numpy(), T_len=T_len.numpy().astype(np.int32),
        W=joint.joint_ln.weight.detach().numpy(), b=joint.joint_ln.bias.detach().numpy(),
        tok_flat=flat, tok_len=np.asarray([len(t) for t in tokens], np.int32),
        max_length=np.int32(max_length), **sd)


if __name__ == "__main__":
    torch.set_num_threads(4)
    # tiny ragged case with U_b = 0 and T_b = 1 rows (dense logits + logit grads kept)
    loss_case("loss_tiny", B=4, T=7, U=4, H=16, V=11, T_len=[7, 5, 1, 3], U_len=[4, 0, 2, 4], seed=11)
    # mid case: kernel-tile friendly sizes, ragged
    loss_case("loss_mid", B=3, T=37, U=13, H=128, V=256, T_len=[37, 20, 31], U_len=[13, 9, 4], seed=12)
    # wide case (H=V=512; the full H=V=1024 width is checked against the live oracle), short lattice
    loss_case("loss_wide", B=2, T=19, U=9, H=512, V=512, T_len=[19, 11], U_len=[7, 9], seed=13)
    decode_case("decode_small", n_utt=6, T=24, H=64, V=48, E=32, seed=21, max_length=40, scale=3.0, blank_bias=2.2)
