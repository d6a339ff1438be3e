# 1 -> 8 GPU ladder on one 8-GPU box (gpurun --gpus 8): the driver's own launch line per N; JSON lines land in gpurun_out/
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extra --sustain-s 0 > gpurun_out/scale_f_n$n.json 2> gpurun_out/scale_f_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_f_n$n.json 2> gpurun_out/scale_f_n$n.err; fi
done
