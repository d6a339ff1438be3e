"""CPU oracle (TEST INFRASTRUCTURE, not product code) for the RNN-T joint + transducer loss path.

fp64 numpy restatement of the reference's algorithm.  Each function cites the reference
file:line it follows (paths relative to /root/reference).  The loss itself lives in a third
party dependency of the reference, torchaudio (2.11.0+cu128 in this image; op
`torchaudio::rnnt_loss_forward`, reference call site rnnt/model.py:35-41); its published
algorithm (Graves 2012 transducer forward/backward in log space, as implemented in
torchaudio's src/libtorchaudio/rnnt/cpu/cpu_kernels.h) is restated here.

Parity pinning: the reference's own tests hold NO golden vector for this path (SURVEY.md
section 4 / 8c), so the oracle is pinned against outputs of the reference itself run in the build
container: tests/golden/make_golden.py imports rnnt.joint.JointNetwork, rnnt.model.RNNTModel
and rnnt.predictor.ConvPredictor from /root/reference plus torchaudio.functional.rnnt_loss,
and the committed fixtures under tests/golden/*.npz are checked by tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- joint
def joint_hidden(enc, pred, audio_ln=None, text_ln=None):
    """tanh(audio.unsqueeze(2) + text.unsqueeze(1))  -- rnnt/joint.py:26-37.

    enc (B,T,Fa), pred (B,U1,Ft); optional (W,b) pre-projections (joint.py:26-30).
    """
    enc = np.asarray(enc, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    if audio_ln is not None:
        enc = enc @ np.asarray(audio_ln[0], np.float64).T + np.asarray(audio_ln[1], np.float64)
    if text_ln is not None:
        pred = pred @ np.asarray(text_ln[0], np.float64).T + np.asarray(text_ln[1], np.float64)
    return np.tanh(enc[:, :, None, :] + pred[:, None, :, :])


def joint_logits(enc, pred, W, b, audio_ln=None, text_ln=None):
    """joint_ln(tanh(a + p))  -- rnnt/joint.py:25-39.  Returns (B,T,U1,V) fp64."""
    h = joint_hidden(enc, pred, audio_ln, text_ln)
    return h @ np.asarray(W, np.float64).T + np.asarray(b, np.float64)


def single_forward(a, p, W, b):
    """rnnt/joint.py:44-55 without broadcasting (decode step)."""
    h = np.tanh(np.asarray(a, np.float64) + np.asarray(p, np.float64))
    return h @ np.asarray(W, np.float64).T + np.asarray(b, np.float64)


# ----------------------------------------------------------------------------- loss pieces
def _logsumexp(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True))).squeeze(axis)


def _lse2(a, b):
    m = np.maximum(a, b)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    return np.where(np.isfinite(m), m_safe + np.log(np.exp(a - m_safe) + np.exp(b - m_safe)), -np.inf)


def log_probs(logits, targets, blank):
    """lse, lpB (skip), lpE (emit)  -- torchaudio ComputeLogProbs, called from rnnt/model.py:35.

    logits (B,T,U1,V); targets (B,U) ints.  lpE[:, :, U] is unused (set to -inf).
    """
    logits = np.asarray(logits, np.float64)
    B, T, U1, V = logits.shape
    if blank < 0:
        blank = V + blank
    lse = _logsumexp(logits, -1)
    lpB = logits[..., blank] - lse
    lpE = np.full((B, T, U1), -np.inf)
    tg = np.asarray(targets).astype(np.int64)
    for b in range(B):
        for u in range(U1 - 1):
            lpE[b, :, u] = logits[b, :, u, tg[b, u]] - lse[b, :, u]
    return lse, lpB, lpE


def lattice(lpB, lpE, T_len, U_len):
    """alpha/beta recursions and costs -- torchaudio ComputeAlphasBetasCosts (model.py:35-41).

    alpha[0,0]=0; alpha[t,u]=LSE(alpha[t-1,u]+lpB[t-1,u], alpha[t,u-1]+lpE[t,u-1])
    beta[Tb-1,Ub]=lpB[Tb-1,Ub]; beta[t,u]=LSE(beta[t+1,u]+lpB[t,u], beta[t,u+1]+lpE[t,u])
    cost_b = -beta[0,0].  Cells outside (Tb, Ub+1) are left at -inf.
    """
    B, T, U1 = lpB.shape
    alpha = np.full((B, T, U1), -np.inf)
    beta = np.full((B, T, U1), -np.inf)
    costs = np.zeros(B)
    for b in range(B):
        Tb, Ub = int(T_len[b]), int(U_len[b])
        a = alpha[b]
        a[0, 0] = 0.0
        for t in range(1, Tb):
            a[t, 0] = a[t - 1, 0] + lpB[b, t - 1, 0]
        for u in range(1, Ub + 1):
            a[0, u] = a[0, u - 1] + lpE[b, 0, u - 1]
        for t in range(1, Tb):
            for u in range(1, Ub + 1):
                a[t, u] = _lse2(a[t - 1, u] + lpB[b, t - 1, u], a[t, u - 1] + lpE[b, t, u - 1])
        be = beta[b]
        be[Tb - 1, Ub] = lpB[b, Tb - 1, Ub]
        for t in range(Tb - 2, -1, -1):
            be[t, Ub] = be[t + 1, Ub] + lpB[b, t, Ub]
        for u in range(Ub - 1, -1, -1):
            be[Tb - 1, u] = be[Tb - 1, u + 1] + lpE[b, Tb - 1, u]
        for t in range(Tb - 2, -1, -1):
            for u in range(Ub - 1, -1, -1):
                be[t, u] = _lse2(be[t + 1, u] + lpB[b, t, u], be[t, u + 1] + lpE[b, t, u])
        costs[b] = -be[0, 0]
    return alpha, beta, costs


def lattice_fast(lpB, lpE, T_len, U_len):
    """Same recursions as `lattice`, vectorised over anti-diagonals (for larger shapes)."""
    B, T, U1 = lpB.shape
    alpha = np.full((B, T, U1), -np.inf)
    beta = np.full((B, T, U1), -np.inf)
    costs = np.zeros(B)
    for b in range(B):
        Tb, Ub = int(T_len[b]), int(U_len[b])
        a = alpha[b]
        a[0, 0] = 0.0
        for n in range(1, Tb + Ub):
            u = np.arange(max(0, n - Tb + 1), min(n, Ub) + 1)
            t = n - u
            up = np.where(t > 0, a[np.maximum(t - 1, 0), u] + lpB[b, np.maximum(t - 1, 0), u], -np.inf)
            lf = np.where(u > 0, a[t, np.maximum(u - 1, 0)] + lpE[b, t, np.maximum(u - 1, 0)], -np.inf)
            a[t, u] = _lse2(up, lf)
        be = beta[b]
        be[Tb - 1, Ub] = lpB[b, Tb - 1, Ub]
        for n in range(Tb + Ub - 2, -1, -1):
            u = np.arange(max(0, n - Tb + 1), min(n, Ub) + 1)
            t = n - u
            dn = np.where(t < Tb - 1, be[np.minimum(t + 1, Tb - 1), u] + lpB[b, t, u], -np.inf)
            rt = np.where(u < Ub, be[t, np.minimum(u + 1, Ub)] + lpE[b, t, u], -np.inf)
            be[t, u] = _lse2(dn, rt)
        costs[b] = -be[0, 0]
    return alpha, beta, costs


def logit_grads(logits, targets, T_len, U_len, blank, lse, lpB, lpE, alpha, beta, dcost=None):
    """d cost_b / d logits -- torchaudio ComputeGradients (computed inside forward, model.py:35).

    g[t,u,v] = softmax_v*gamma - [v=blank]*eB - [u<Ub and v=y_u]*eE, gamma = exp(alpha+beta-logZ),
    eB = exp(alpha+lpB+beta[t+1,u]-logZ) (terminal cell uses beta==0, last row u<Ub has none),
    eE = exp(alpha+lpE+beta[t,u+1]-logZ).  Exactly 0 on padding.  clamp<=0 -> no clamp.
    """
    logits = np.asarray(logits, np.float64)
    B, T, U1, V = logits.shape
    if blank < 0:
        blank = V + blank
    g = np.zeros_like(logits)
    tg = np.asarray(targets).astype(np.int64)
    for b in range(B):
        Tb, Ub = int(T_len[b]), int(U_len[b])
        logZ = beta[b, 0, 0]
        dc = 1.0 if dcost is None else float(dcost[b])
        for t in range(Tb):
            for u in range(Ub + 1):
                c = alpha[b, t, u] - logZ
                gam = np.exp(c + beta[b, t, u])
                row = np.exp(logits[b, t, u] - lse[b, t, u]) * gam
                if t < Tb - 1:
                    row[blank] -= np.exp(c + lpB[b, t, u] + beta[b, t + 1, u])
                elif u == Ub:
                    row[blank] -= np.exp(c + lpB[b, t, u])
                if u < Ub:
                    row[tg[b, u]] -= np.exp(c + lpE[b, t, u] + beta[b, t, u + 1])
                g[b, t, u] = row * dc
    return g


def loss_and_grads(enc, pred, W, b, targets, T_len, U_len, blank=-1, dcost=None, fast=True, audio_ln=None,
                   text_ln=None):
    """Full path rnnt/joint.py:25-39 -> rnnt/model.py:35-41 -> autograd (train.py:134).

    Returns dict(costs[B], d_enc, d_pred, dW, db, lse, lpB, lpE, alpha, beta) in fp64.
    dcost defaults to ones (reduction='none' summed); pass 1/B for reduction='mean'.
    audio_ln / text_ln = (weight, bias) of the optional pre-projections (joint.py:8-12,26-30): enc / pred are then
    the RAW inputs, d_enc / d_pred their gradients, and dWa, dba, dWt, dbt are added to the result.
    """
    enc = np.asarray(enc, np.float64); pred = np.asarray(pred, np.float64)
    W = np.asarray(W, np.float64); b = np.asarray(b, np.float64)
    h = joint_hidden(enc, pred, audio_ln, text_ln)
    logits = h @ W.T + b
    lse, lpB, lpE = log_probs(logits, targets, blank)
    alpha, beta, costs = (lattice_fast if fast else lattice)(lpB, lpE, T_len, U_len)
    g = logit_grads(logits, targets, T_len, U_len, blank, lse, lpB, lpE, alpha, beta, dcost)
    B, T, U1, V = logits.shape
    g2 = g.reshape(-1, V); h2 = h.reshape(-1, h.shape[-1])
    dW = g2.T @ h2                      # autograd of joint_ln, joint.py:39
    db = g2.sum(0)
    dz = (g2 @ W).reshape(h.shape) * (1.0 - h * h)   # tanh backward, joint.py:37
    out = dict(costs=costs, d_enc=dz.sum(2), d_pred=dz.sum(1), dW=dW, db=db,
               lse=lse, lpB=lpB, lpE=lpE, alpha=alpha, beta=beta)
    if audio_ln is not None:            # autograd of audio_ln, joint.py:26-27
        Wa = np.asarray(audio_ln[0], np.float64)
        da = out["d_enc"]
        out["dWa"] = da.reshape(-1, da.shape[-1]).T @ enc.reshape(-1, enc.shape[-1])
        out["dba"] = da.reshape(-1, da.shape[-1]).sum(0)
        out["d_enc"] = da @ Wa
    if text_ln is not None:             # autograd of text_ln, joint.py:29-30
        Wt = np.asarray(text_ln[0], np.float64)
        dt = out["d_pred"]
        out["dWt"] = dt.reshape(-1, dt.shape[-1]).T @ pred.reshape(-1, pred.shape[-1])
        out["dbt"] = dt.reshape(-1, dt.shape[-1]).sum(0)
        out["d_pred"] = dt @ Wt
    return out


# ----------------------------------------------------------------------------- predictor + decode
def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _gelu(x):
    from math import sqrt
    try:
        from scipy.special import erf
    except Exception:  # pragma: no cover
        erf = np.vectorize(__import__("math").erf)
    return 0.5 * x * (1.0 + erf(x / sqrt(2.0)))


def _causal_conv(x, w, b):
    """rnnt/causalconv.py:23-30 with stride 1, dilation 1: left zero pad k-1.  x (L,C), w (Co,Ci,k)."""
    k = w.shape[2]
    L = x.shape[0]
    xp = np.concatenate([np.zeros((k - 1, x.shape[1])), x], 0)
    out = np.zeros((L, w.shape[0]))
    for j in range(k):
        out += xp[j:j + L] @ w[:, :, j].T
    return out + b


def conv_predictor(ids, sd):
    """ConvPredictor.forward in eval mode (dropout off) -- rnnt/predictor.py:209-229.  ids (L,) -> (L,D)."""
    f = lambda k: np.asarray(sd[k], np.float64)
    x = f("embedding.weight")[np.asarray(ids, np.int64)]
    x = _layer_norm(x, f("input_layer_norm.weight"), f("input_layer_norm.bias"))
    x = _gelu(_causal_conv(x, f("conv1.conv.weight"), f("conv1.conv.bias")))
    x = _gelu(_causal_conv(x, f("conv2.conv.weight"), f("conv2.conv.bias")))
    x = x @ f("linear.weight").T + f("linear.bias")
    return _layer_norm(x, f("output_layer_norm.weight"), f("output_layer_norm.bias"))


def greedy_decode(enc, T_b, W, b, pred_sd, blank, max_length=200, max_per_frame=10):
    """RNNTModel._greedy_decode_conv on given encoder features -- rnnt/model.py:90-128.

    enc (T,H) for ONE utterance.  Returns (tokens, margins): margins = top1-top2 logit gap per step.
    """
    tokens = [blank]
    t = 0
    per = 0
    feats = conv_predictor(tokens, pred_sd)
    margins = []
    while t < T_b and len(tokens) < max_length:
        logits = single_forward(enc[t], feats[-1], W, b)
        tok = int(np.argmax(logits))
        srt = np.sort(logits)
        margins.append(float(srt[-1] - srt[-2]))
        if tok == blank or per >= max_per_frame:
            t += 1
            per = 0
        else:
            tokens.append(tok)
            feats = conv_predictor(tokens, pred_sd)
            per += 1
    return tokens[1:], margins
