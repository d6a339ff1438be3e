"""Opcode census of the built library: per kernel, how many tcgen05 / TMEM / TMA / legacy-MMA instructions the SASS holds.
    python scripts/sass_census.py > profiles/r2_sass_census.txt
SASS mnemonics (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG /
UTMASTG, tcgen05.commit -> UTCBAR, mma.sync -> HMMA (must be absent), cp.async -> LDGSTS."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rnnt_b200", "_C", "librnnt_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR",
            "SYNCS", "HMMA", "HGMMA", "LDGSTS", "MUFU.TANH", "MUFU.EX2", "MUFU.LG2", "RED.E", "REDG", "ATOMG", "ATOM",
            "FFMA2", "FADD2", "FMUL2", "FMNMX3", "FFMA", "BAR.SYNC", "UCGABAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    kernels[cur][p] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and ".2CTA" in op:
                kernels[cur]["UTMALDG.2CTA"] += 1
            if op.startswith("UTCBAR") and "MULTICAST" in op:
                kernels[cur]["UTCBAR.MULTICAST"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    tot = collections.Counter()
    for (name, cnt), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "")).replace("rb::", "").replace("void ", "")
        keys = [k for k in cnt if k != "_total" and cnt[k]]
        print(f"{short:58s} instrs {cnt['_total']:6d}  " + "  ".join(f"{k}={cnt[k]}" for k in sorted(keys)))
        tot.update(cnt)
    print("# totals: " + "  ".join(f"{k}={tot[k]}" for k in sorted(tot) if k != "_total"))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core path found"


if __name__ == "__main__":
    sys.exit(main())
