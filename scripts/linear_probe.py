"""Rate of the generic tcgen05 GEMM (linear_gemm.cu) per operand layout: K-major x K-major (forward), K-major x MN-major
(d-input), MN-major x MN-major (d-weight, no split needed at this size).  Includes the fp32->fp16 operand conversion."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnnt_b200 import _lib
L = _lib.lib()
M, K, N = 8192, 4096, 4096
x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / 64; b = torch.zeros(N, device="cuda")
dy = torch.randn(M, N, device="cuda")
y = torch.empty(M, N, device="cuda"); dx = torch.empty(M, K, device="cuda"); dW = torch.empty(N, K, device="cuda")
nb = L.rnnt_b200_linear_workspace_bytes(M, K, N, 1, 0)
ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fl = 2.0 * M * K * N
f = lambda: L.rnnt_b200_linear_fwd(x.data_ptr(), W.data_ptr(), b.data_ptr(), M, K, N, y.data_ptr(), ws.data_ptr(), nb, st)
g = lambda: L.rnnt_b200_linear_bwd(x.data_ptr(), W.data_ptr(), dy.data_ptr(), M, K, N, dx.data_ptr(), None, None, 0, ws.data_ptr(), nb, st)
h = lambda: L.rnnt_b200_linear_bwd(x.data_ptr(), W.data_ptr(), dy.data_ptr(), M, K, N, None, dW.data_ptr(), None, 0, ws.data_ptr(), nb, st)
c = lambda: (x.half(), W.half())
tc = t(c)
for name, fn in (("fwd  K-major x K-major ", f), ("dx   K-major x MN-major", g), ("dW   MN-major x MN-major", h)):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms  ({fl/ms/1e9:.0f} TF/s incl. conversions; two torch .half() conversions of this size take {tc:.3f} ms)")
ms = t(lambda: torch.matmul(x.half(), W.half().t()))
print(f"torch fp16 matmul incl. conversions: {ms:.3f} ms ({fl/ms/1e9:.0f} TF/s)")
