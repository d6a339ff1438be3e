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


def ref_conv_predictor(ids, sd):
    """rnnt/predictor.py:209-229 (eval mode) with torch functional ops on a state_dict.  ids (1,L) int64 -> (1,L,D)."""
    F = torch.nn.functional
    x = F.embedding(ids, sd["embedding.weight"])
    x = F.layer_norm(x, x.shape[-1:], sd["input_layer_norm.weight"], sd["input_layer_norm.bias"])
    x = x.permute(0, 2, 1)
    for name in ("conv1", "conv2"):                                   # rnnt/causalconv.py:23-30: left zero padding k-1
        w = sd[name + ".conv.weight"]
        x = F.gelu(F.conv1d(F.pad(x, (w.shape[2] - 1, 0)), w, sd[name + ".conv.bias"]))
    x = F.linear(x.permute(0, 2, 1), sd["linear.weight"], sd["linear.bias"])
    return F.layer_norm(x, x.shape[-1:], sd["output_layer_norm.weight"], sd["output_layer_norm.bias"])


@torch.no_grad()
def ref_greedy_decode(audio_features, W, b, pred_sd, blank, max_length=200, max_outputs_per_step=10):
    """rnnt/model.py:90-128 (_greedy_decode_conv) for ONE utterance on given encoder features (1,T,H): joint
    single_forward (joint.py:44-55) + argmax + `.item()` per step, full-history predictor re-run per emitted token."""
    tokens = [blank]
    t, per, T = 0, 0, audio_features.shape[1]
    feats = ref_conv_predictor(torch.tensor([tokens], dtype=torch.int64), pred_sd)
    while t < T and len(tokens) < max_length:
        logits = torch.nn.functional.linear(torch.tanh(audio_features[:, t, :] + feats[:, -1, :]), W, b)
        tok = logits.argmax(dim=-1).item()
        if tok == blank or per >= max_outputs_per_step:
            t += 1
            per = 0
        else:
            tokens.append(tok)
            feats = ref_conv_predictor(torch.tensor([tokens], dtype=torch.int64), pred_sd)
            per += 1
    return tokens[1:]
