"""Format gpurun_out/scale_<tag>_n{1,2,4,8}.json (scripts/scaling_table.sh) into the table kept under profiles/."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "f"
base = None
print("# bench.py --gpus N --steps 20 --warmup 3 on one 8-GPU B200 box (gpurun --gpus 8), final round-2 kernels")
print("# N>1: BASELINE configs[2], global batch 256 sharded by utterance (strong); weak = B=32 per GPU (round-1 workload)")
for n in (1, 2, 4, 8):
    d = json.loads(open(f"gpurun_out/scale_{tag}_n{n}.json").read().strip().splitlines()[-1])
    v = d["value"]
    base = base or v
    w = d.get("weak_b32_per_gpu") or {}
    nccl = (d.get("nccl") or {}).get("exposed_ms_per_step", 0.0)
    print(f"N={n}  {d['scaling']:6s} B/GPU={d['config']['per_gpu_batch']:3d}  ms/step {d['ms_per_step']:7.3f}  value {v/1e6:8.1f} M cells/s  "
          f"x{v/base:5.2f} (eff {v/base/n:.3f})  e2e {d['e2e']['value']/1e6:8.1f} M  nccl exposed {nccl:.3f} ms  "
          f"weak(B=32/GPU) {w.get('value', 0)/1e6:8.1f} M x{w.get('value', 0)/base:5.2f}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
