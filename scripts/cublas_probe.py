"""Reference point for the roofline: one cuBLAS bf16 GEMM (8192^3, and the joint's own 1.3M x 1024 x 1024 shape)."""
import sys, torch
torch.manual_seed(0)
shapes = [(8192, 8192, 8192), (32 * 400 * 101, 1024, 1024)]
for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        c = a @ b.t()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        c = a @ b.t()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"cuBLAS bf16 {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
    del a, b, c
