"""Pin the CPU oracle against golden vectors produced by the reference itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import rnnt_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["loss_tiny", "loss_mid", "loss_wide"])
def test_oracle_costs_and_grads_match_reference(golden_dir, name):
    g = _load(golden_dir, name)
    out = orc.loss_and_grads(g["enc"], g["pred"], g["W"], g["b"], g["targets"], g["T_len"], g["U_len"], blank=-1)
    np.testing.assert_allclose(out["costs"], g["costs"], rtol=2e-6, atol=2e-5)
    assert abs(out["costs"].mean() - float(g["loss_mean"])) < 1e-4 * abs(float(g["loss_mean"]))
    for k in ("d_enc", "d_pred", "dW", "db"):
        ref = g[k]
        err = np.abs(out[k] - ref).max()
        assert err <= 1e-4 * max(1.0, np.abs(ref).max()), (k, err)   # fp32 reference rounding


def test_oracle_slow_and_fast_lattice_agree(golden_dir):
    g = _load(golden_dir, "loss_tiny")
    logits = orc.joint_logits(g["enc"], g["pred"], g["W"], g["b"])
    lse, lpB, lpE = orc.log_probs(logits, g["targets"], -1)
    a1, b1, c1 = orc.lattice(lpB, lpE, g["T_len"], g["U_len"])
    a2, b2, c2 = orc.lattice_fast(lpB, lpE, g["T_len"], g["U_len"])
    np.testing.assert_allclose(c1, c2, rtol=1e-12)
    m = np.isfinite(a1)
    np.testing.assert_allclose(a1[m], a2[m], rtol=1e-12, atol=1e-12)
    m = np.isfinite(b1)
    np.testing.assert_allclose(b1[m], b2[m], rtol=1e-12, atol=1e-12)


def test_oracle_logits_and_logit_grads(golden_dir):
    g = _load(golden_dir, "loss_tiny")
    logits = orc.joint_logits(g["enc"], g["pred"], g["W"], g["b"])
    np.testing.assert_allclose(logits, g["logits"], rtol=0, atol=5e-6)
    lse, lpB, lpE = orc.log_probs(logits, g["targets"], -1)
    al, be, _ = orc.lattice(lpB, lpE, g["T_len"], g["U_len"])
    dl = orc.logit_grads(logits, g["targets"], g["T_len"], g["U_len"], -1, lse, lpB, lpE, al, be)
    np.testing.assert_allclose(dl, g["dlogits"], rtol=0, atol=3e-6)
    # padded cells: exactly zero gradient; valid cells: rows sum to ~0
    B, T, U1, V = dl.shape
    for b in range(B):
        Tb, Ub = int(g["T_len"][b]), int(g["U_len"][b])
        assert not dl[b, Tb:].any() and not dl[b, :, Ub + 1:].any()
        assert np.abs(dl[b, :Tb, :Ub + 1].sum(-1)).max() < 1e-12


def test_oracle_known_answers():
    rng = np.random.default_rng(0)
    # U_b = 0: the only path is all blanks -> cost = -sum_t lpB[t,0]
    T, V = 5, 7
    logits = rng.normal(size=(1, T, 1, V))
    lse, lpB, lpE = orc.log_probs(logits, np.zeros((1, 0), np.int32), -1)
    _, _, c = orc.lattice(lpB, lpE, [T], [0])
    assert abs(c[0] + lpB[0, :, 0].sum()) < 1e-12
    # T_b = 1: emit all U labels then one blank
    U = 3
    logits = rng.normal(size=(1, 1, U + 1, V))
    tg = rng.integers(0, V - 1, size=(1, U))
    lse, lpB, lpE = orc.log_probs(logits, tg, -1)
    _, _, c = orc.lattice(lpB, lpE, [1], [U])
    assert abs(c[0] + lpE[0, 0, :U].sum() + lpB[0, 0, U]) < 1e-12


def test_oracle_greedy_decode_matches_reference(golden_dir):
    g = _load(golden_dir, "decode_small")
    sd = {k[5:]: g[k] for k in g.files if k.startswith("pred.")}
    V = g["W"].shape[0]
    off = 0
    for i, n in enumerate(g["tok_len"]):
        want = g["tok_flat"][off:off + n].tolist()
        off += n
        got, margins = orc.greedy_decode(g["feats"][i], int(g["T_len"][i]), g["W"], g["b"], sd, blank=V - 1,
                                         max_length=int(g["max_length"]))
        assert got == want, (i, min(margins))
