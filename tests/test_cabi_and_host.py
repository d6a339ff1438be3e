"""CPU-only tests: the C-ABI library loads and exports what include/rnnt_b200.h declares (no compute calls), the
host-side logic (chunking, sharding, module surface, error paths), and the N>1 path over gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rnnt_b200.h")).read()
    return sorted(set(re.findall(r"\b(rnnt_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rnnt_b200 import _lib
    from rnnt_b200.build import build_extension
    build_extension()
    handle = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(handle, s), s
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS)
    assert handle.rnnt_b200_abi_version() == 3


def test_ctypes_signatures_match_the_header_prototypes():
    """Every binding in rnnt_b200/_lib.py passes as many arguments as the prototype in include/rnnt_b200.h declares."""
    from rnnt_b200 import _lib
    text = open(os.path.join(ROOT, "include", "rnnt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = dict(re.findall(r"\b(rnnt_b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))
    assert set(protos) == set(_lib._SIGNATURES)
    for name, params in protos.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib._SIGNATURES[name][1]), (name, n, len(_lib._SIGNATURES[name][1]))


def test_header_is_plain_c(tmp_path):
    """include/rnnt_b200.h is the drop-in boundary: it must compile as C99 with nothing but <stddef.h> / <stdint.h>, and a
    C translation unit that takes the address of every declared entry point must link against the library."""
    syms = _header_symbols()
    src = tmp_path / "abi_check.c"
    src.write_text('#include "rnnt_b200.h"\n#include <stdio.h>\nint main(void) {\n  const void* fns[] = {\n'
                   + "".join(f"    (const void*)&{s},\n" for s in syms)
                   + '  };\n  printf("%d %d\\n", (int)(sizeof(fns) / sizeof(fns[0])), rnnt_b200_abi_version());\n  return 0;\n}\n')
    from rnnt_b200.build import LIB_PATH, build_extension
    build_extension()
    exe = tmp_path / "abi_check"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", os.path.dirname(LIB_PATH), "-lrnnt_b200", "-Wl,-rpath," + os.path.dirname(LIB_PATH)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == [str(len(syms)), "3"], (r.stdout, r.stderr)


def test_argument_validation_without_gpu():
    from rnnt_b200 import _lib
    L = _lib.lib()
    assert L.rnnt_b200_max_tiles(32, 400, 101) == 32 * 25 * 13
    fwd, bwd = C.c_size_t(0), C.c_size_t(0)
    assert L.rnnt_b200_workspace_bytes(32, 400, 101, 1024, 1024, 2080, 0, 0, C.byref(fwd), C.byref(bwd)) == 0
    assert fwd.value >= 1024 * 1024 * 2 and bwd.value >= 2080 * 128 * 2048 * 2
    assert L.rnnt_b200_workspace_bytes(0, 400, 101, 1024, 1024, 1, 0, 0, C.byref(fwd), C.byref(bwd)) < 0
    assert b"invalid shape" in L.rnnt_b200_last_error()
    # invalid arguments are rejected before anything touches a device
    rc = L.rnnt_b200_joint_loss_fwd(None, 0, 0, 1, None, None, None, None, None, None, 2, 8, 3, 12, 16, -1,
                                    None, None, None, None, None, None, None, None, 0, None)
    assert rc == -2 and b"multiple of 8" in L.rnnt_b200_last_error()
    rc = L.rnnt_b200_joint_loss_fwd(None, 0, 0, 1, None, None, None, None, None, None, 2, 8, 3000, 16, 16, -1,
                                    None, None, None, None, None, None, None, None, 0, None)
    assert L.rnnt_b200_hidden_bytes(32, 400, 101, 1024) == 32 * 25 * 13 * 128 * 1024 * 2
    assert rc == -5


def test_product_path_refuses_cpu_tensors():
    import rnnt_b200
    enc, pred = torch.randn(1, 4, 16), torch.randn(1, 3, 16)
    W, b = torch.randn(8, 16), torch.randn(8)
    tg = torch.zeros(1, 2, dtype=torch.int32)
    tl, ul = torch.tensor([4], dtype=torch.int32), torch.tensor([2], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rnnt_b200.joint_rnnt_loss(enc, pred, W, b, tg, tl, ul)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rnnt_b200.rnnt_loss(torch.randn(1, 4, 3, 8), tg, tl, ul)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rnnt_b200.joint_argmax(enc[0], enc[0], W, b)


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rnnt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_joint_module_surface_matches_reference():
    import rnnt_b200
    j = rnnt_b200.JointNetwork(-1, -1, 32, 11)
    assert j.blank_idx == 10 and sorted(j.state_dict()) == ["joint_ln.bias", "joint_ln.weight"]
    assert tuple(j.joint_ln.weight.shape) == (11, 32)
    a, t = torch.randn(2, 5, 32), torch.randn(2, 3, 32)
    out = j(a, t)
    assert out.shape == (2, 5, 3, 11)
    torch.testing.assert_close(out, torch.nn.functional.linear(torch.tanh(a.unsqueeze(2) + t.unsqueeze(1)),
                                                               j.joint_ln.weight, j.joint_ln.bias))
    assert j.single_forward(a[:, 0], t[:, 0]).shape == (2, 11)
    j2 = rnnt_b200.JointNetwork(7, 9, 32, 11)
    assert j2(torch.randn(2, 5, 7), torch.randn(2, 3, 9)).shape == (2, 5, 3, 11)


def test_conv_predictor_mirror_and_window_equivalence():
    import rnnt_b200
    torch.manual_seed(0)
    p = rnnt_b200.ConvPredictor(50, 24, 16, 0.3).eval()
    assert set(p.state_dict()) == {
        "embedding.weight", "input_layer_norm.weight", "input_layer_norm.bias", "conv1.conv.weight",
        "conv1.conv.bias", "conv2.conv.weight", "conv2.conv.bias", "linear.weight", "linear.bias",
        "output_layer_norm.weight", "output_layer_norm.bias"}
    ids = torch.randint(0, 50, (1, 15))
    full = p(ids)
    for n in range(1, 16):
        win = ids[:, max(0, n - 7):n]
        torch.testing.assert_close(p.last_step(win), full[:, n - 1], rtol=1e-5, atol=1e-6)


def test_incremental_predictor_equals_full_rerun():
    """ConvPredictorStepper (device-side decode state) reproduces the reference's full-history re-run exactly."""
    import rnnt_b200
    from rnnt_b200.predictor import ConvPredictorStepper
    torch.manual_seed(1)
    p = rnnt_b200.ConvPredictor(40, 24, 16, 0.3).eval()
    ids = torch.randint(0, 40, (4, 11))
    full = p(ids)
    st = ConvPredictorStepper(p, 4, "cpu")
    everyone = torch.ones(4, dtype=torch.bool)
    for i in range(11):
        torch.testing.assert_close(st.advance(ids[:, i], everyone), full[:, i], rtol=1e-5, atol=1e-5)
    # rows that do not emit keep their state
    st2 = ConvPredictorStepper(p, 4, "cpu")
    mask = torch.tensor([True, False, True, False])
    st2.advance(ids[:, 0], everyone)
    st2.advance(ids[:, 1], mask)
    out = st2.advance(ids[:, 2], everyone)
    torch.testing.assert_close(out[0], full[0, 2], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[1], p(ids[1:2, [0, 2]])[0, 1], rtol=1e-5, atol=1e-5)


def test_ring_chunking_and_sharding_helpers():
    from rnnt_b200.functional import pick_ring_tiles
    from rnnt_b200.parallel import balanced_assignment, shard_bounds
    rt = pick_ring_tiles(32, 400, 101, 1024, 1024, ring_bytes=1 << 30)
    assert rt * 128 * 4096 <= (1 << 30) and rt * 6 >= 10400 > rt * 4
    assert pick_ring_tiles(2, 20, 10, 64, 256) == 2 * 2 * 2
    assert shard_bounds(256, 8, 3) == (96, 128)
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
    buckets, loads = balanced_assignment([100, 90, 50, 40, 30, 10], 2)
    assert sorted(sum(buckets, [])) == list(range(6)) and abs(loads[0] - loads[1]) <= 20


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rnnt_b200.parallel import GradAllReducer, WeightGradBucket, shard_bounds
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_bounds(8, world, rank)
assert (lo, hi) == (rank * 4, rank * 4 + 4)
g1 = torch.full((5, 3), float(rank + 1)); g2 = torch.arange(4, dtype=torch.float32) * (rank + 1)
red = GradAllReducer([], average=True)
red.all_reduce_grads([g1, g2])
assert torch.allclose(g1, torch.full((5, 3), 1.5)) and torch.allclose(g2, torch.arange(4) * 1.5), (g1, g2)
red_sum = GradAllReducer([], average=False)
g3 = torch.ones(7) * (rank + 1)
red_sum.all_reduce_grads([g3])
assert torch.allclose(g3, torch.full((7,), 3.0))
# flat bucket of the overlapped path: the sink protocol the fused backward drives, on CPU tensors
bk = WeightGradBucket(6, 4, 5, "cpu", average=True)
assert bk.accepts(6, 4, "cpu") and not bk.accepts(6, 8, "cpu")
dW, db, ev = bk.weight_grad_views(6, 4)
assert ev is None and dW.shape == (6, 4) and db.shape == (6,) and dW.data_ptr() == bk.flat.data_ptr()
dW.fill_(float(rank + 1)); db.fill_(2.0 * (rank + 1)); bk.extra_slice.fill_(4.0 * (rank + 1))
bk.weight_grads_enqueued()           # joint slice reduced early ...
flat = bk.finish()                   # ... the rest + averaging at the end of the step
assert torch.allclose(flat[:24], torch.full((24,), 1.5)) and torch.allclose(flat[24:30], torch.full((6,), 3.0))
assert torch.allclose(flat[30:], torch.full((5,), 6.0)) and not bk.pending
dW.fill_(float(rank + 1)); db.fill_(0.0); bk.extra_slice.fill_(0.0)
flat = bk.finish()                   # no fused backward fed the bucket: finish() reduces everything
assert torch.allclose(flat[:24], torch.full((24,), 1.5))
dist.destroy_process_group()
print("ok", rank)
"""


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_bench_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    code = ("import bench, json; bench.CPU_SAMPLE.update(B=1, T=24, U=6); "
            "import argparse; bench.run_reference_arm(argparse.Namespace(gpus=1, steps=1, warmup=1))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "lattice-cells/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
