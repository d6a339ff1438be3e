"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <report.ncu-rep> <launches.csv> <tag>      (runs here, no GPU needed)

Writes profiles/<tag>_ncu_full_summary.txt (key metrics per captured launch), profiles/<tag>_traffic.json
(DRAM bytes per launch per kernel family, read by bench.py for roofline.traffic) and
profiles/<tag>_launches_summary.txt (per-kernel totals / shares of the launch list).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "smsp__inst_executed_pipe_lsu.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__cluster_dim_x", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def family(name):
    if "joint_gemm_kernel<0" in name: return "joint_gemm_fwd"
    if "joint_gemm_kernel<1" in name: return "joint_gemm_bwd"
    if "dh_gemm" in name: return "dh_gemm"
    if "dw_gemm" in name: return "dw_gemm"
    if "lattice_kernel" in name: return "lattice"
    return None


def main():
    rep, launches, tag = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = [f"# source: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on, one bench.py step:",
           "# forward, lattice, then per ring chunk recompute/gradient kernel, dh GEMM, dW GEMM); the report itself is scratch",
           "# (gpurun_out/) and not committed.  Times under ncu are serialised and cold-cache: use the SHARES.", ""]
    traffic = collections.defaultdict(lambda: dict(bytes=0.0, n=0))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        out.append(name[:110])
        for m in METRICS:
            if m in col:
                out.append(f"  {m:<72s} {r[col[m]]:>18s} {units[col[m]]}")
        out.append("")
        fam = family(name)
        dur = float(r[col["gpu__time_duration.sum"]])
        if fam and dur > 0.02:      # skip the empty-chunk launches
            b = sum(float(r[col[k]]) * UNIT_SCALE.get(units[col[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic[fam]["bytes"] += b
            traffic[fam]["n"] += 1
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt"), "w") as f:
        f.write("\n".join(out))
    tj = {k: dict(dram_bytes_per_launch=v["bytes"] / v["n"], launches_sampled=v["n"],
                  source=f"profiles/{tag}_ncu_full_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)")
          for k, v in traffic.items()}
    with open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    # launch list
    text = [l for l in open(launches) if not l.startswith("==")]
    rd = list(csv.DictReader(text))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "ns")
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        k = r["Kernel Name"][:64]
        tot[k][0] += 1
        tot[k][1] += us
    total = sum(v[1] for v in tot.values())
    lines = [f"# launch list summary: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --sustain-s 0",
             "# per-launch times are cold-cache and serialised; compare SHARES with bench.py's live CUDA-event shares",
             f"{'kernel':<66s}{'launches':>9s}{'total_us':>12s}{'share':>8s}"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:<66s}{v[0]:>9d}{v[1]:>12.1f}{100 * v[1] / total:>7.1f}%")
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
