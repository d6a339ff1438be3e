# Round-end capture on one B200: tests, default bench line, reference arm, ncu launch list + full set of one step, decode.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_final.txt
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_final.json 2> gpurun_out/ref_final.err
python scripts/decode_bench.py > gpurun_out/decode_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --sustain-s 0 > gpurun_out/ncu1_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'joint_gemm|dh_gemm|dw_gemm|lattice_kernel' -s 20 -c 5 \
  -o gpurun_out/prof_r2f -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --sustain-s 0 --no-kernel-profile > gpurun_out/ncu2_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:greedy_decode -c 1 -o gpurun_out/prof_decode_r2f -f python scripts/decode_bench.py > gpurun_out/ncu3_final.log 2>&1
ls -la gpurun_out/*.ncu-rep
