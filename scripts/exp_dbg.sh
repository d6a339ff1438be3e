for d in ${DBGS:-0 1 2 3}; do
  RNNT_B200_DBG=$d python bench.py --no-cpu-baseline --no-extra --sustain-s 0 > gpurun_out/x_dbg$d.json 2> gpurun_out/x_dbg$d.err
done
