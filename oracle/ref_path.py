"""TEST INFRASTRUCTURE: the reference's joint + loss call path restated with the same third-party
ops it uses (torch nn.functional.linear + torchaudio.functional.rnnt_loss), runnable on the GPU box
where /root/reference does not exist.  Used by tests as a second checker and by bench.py as the CPU
baseline ("port": same ops, same call arguments as rnnt/joint.py:25-39 and rnnt/model.py:35-41).
"""
from __future__ import annotations

import torch


def ref_joint_forward(enc, pred, W, b):
    """rnnt/joint.py:32-39 (no audio_ln/text_ln: the working configs use features=-1)."""
    joint = torch.tanh(enc.unsqueeze(2) + pred.unsqueeze(1))
    return torch.nn.functional.linear(joint, W, b)


def ref_loss(enc, pred, W, b, targets, T_len, U_len, reduction="mean"):
    """rnnt/model.py:32-41: joint -> torchaudio.functional.rnnt_loss(blank=-1, clamp=-1)."""
    import torchaudio
    logits = ref_joint_forward(enc, pred, W, b)
    return torchaudio.functional.rnnt_loss(
        logits=logits, targets=targets.int(), logit_lengths=T_len.int(),
        target_lengths=U_len.int(), blank=-1, clamp=-1, reduction=reduction)


def ref_loss_and_grads(enc, pred, W, b, targets, T_len, U_len, reduction="none", dcost=None):
    """Returns (costs or loss, d_enc, d_pred, dW, db) through autograd, as train.py:133-134 does."""
    enc = enc.detach().clone().requires_grad_(True)
    pred = pred.detach().clone().requires_grad_(True)
    W = W.detach().clone().requires_grad_(True)
    b = b.detach().clone().requires_grad_(True)
    out = ref_loss(enc, pred, W, b, targets, T_len, U_len, reduction=reduction)
    if out.dim() == 0:
        out.backward()
    else:
        out.backward(torch.ones_like(out) if dcost is None else dcost)
    return out.detach(), enc.grad, pred.grad, W.grad, b.grad
