# same-box A/B: base.so (previous build) vs the current build; DBGS = dbg masks to try with the current build
run() { python bench.py --no-cpu-baseline --no-extra --sustain-s ${SUS:-0} > gpurun_out/x_$1.json 2> gpurun_out/x_$1.err; }
RNNT_B200_LIB=$PWD/rnnt_b200/_C/base.so run base
for d in ${DBGS:-0}; do RNNT_B200_DBG=$d run dbg$d; done
RNNT_B200_LIB=$PWD/rnnt_b200/_C/base.so run base2
