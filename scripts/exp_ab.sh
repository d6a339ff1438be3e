# same-box A/B of library builds: LIBS = names under rnnt_b200/_C (without .so), each timed twice in alternation
run() { python bench.py --no-cpu-baseline --no-extra --sustain-s ${SUS:-0} > gpurun_out/x_$1.json 2> gpurun_out/x_$1.err; }
for rep in 1 2; do for l in ${LIBS:-base}; do RNNT_B200_LIB=$PWD/rnnt_b200/_C/$l.so run ${l}_$rep; done; done
