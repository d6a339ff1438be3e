#!/bin/bash
# A/B of the encoder layouts on one box: headline + sustained, per-kernel times
for i in 1 2; do
python bench.py --no-extra --no-cpu-baseline > gpurun_out/ab_view_$i.json 2>> gpurun_out/ab.err
python bench.py --no-extra --no-cpu-baseline --dense-encoder > gpurun_out/ab_dense_$i.json 2>> gpurun_out/ab.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    d=json.load(open(f)); r=d["roofline"]; k=r["kernels"]; s=r["sustained"]["kernels"]
    print(f, "step %.3f F %.3f G %.3f dh %.3f dW %.3f | sustained step %.3f F %.3f G %.3f dh %.3f dW %.3f all_tiles %.3f" % (
        d["ms_per_step"], k["joint_gemm_fwd"]["ms_per_step"], k["joint_gemm_bwd"]["ms_per_step"], k["dh_gemm"]["ms_per_step"], k["dw_gemm"]["ms_per_step"],
        r["sustained"]["ms_per_step"], s["joint_gemm_fwd"]["ms_per_step"], s["joint_gemm_bwd"]["ms_per_step"], s["dh_gemm"]["ms_per_step"], s["dw_gemm"]["ms_per_step"], d["all_tiles"]["ms_per_step"]))
PY
