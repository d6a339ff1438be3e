"""All-reduce latency of the step's gradient buckets in isolation (torchrun, one rank per GPU)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for n in (1_050_624, 3_150_848, 4_201_472):
    x = torch.ones(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"world {dist.get_world_size()} all_reduce {n*4/1e6:.1f} MB: {e0.elapsed_time(e1)/50*1e3:.1f} us", flush=True)
dist.destroy_process_group()
