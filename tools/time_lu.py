"""LU factor/solve timing at the bench size (tools, not product)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilm_b200 as ilm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4593
torch.manual_seed(0)
A = torch.randn(n, n, dtype=torch.float64, device="cuda") + 0.0 * torch.eye(n, dtype=torch.float64, device="cuda")
for _ in range(2):
    lu = ilm.LU(A)
torch.cuda.synchronize(); t0 = time.perf_counter()
lu = ilm.LU(A); torch.cuda.synchronize()
print("LU ms", (time.perf_counter() - t0) * 1e3)
b = torch.randn(n, dtype=torch.float64, device="cuda")
x = lu.solve(b); torch.cuda.synchronize(); t0 = time.perf_counter()
x = lu.solve(b); torch.cuda.synchronize()
print("solve ms", (time.perf_counter() - t0) * 1e3, "residual", ((A @ x - b).abs().max() / b.abs().max()).item())
