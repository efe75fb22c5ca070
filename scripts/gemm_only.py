"""Run the three tcgen05 GEMMs at the C2 row count a few times (target of the ncu --set full capture)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spatial-vae_b200"))
import torch
import spatial_vae.functional as SF
rows, H, Hp = int(os.environ.get("ROWS", 1024 * 784)), 500, 512
dev = torch.device("cuda:0")
A = (torch.randn(rows, Hp, device=dev) * 0.5).bfloat16()
D = (torch.randn(rows, Hp, device=dev) * 0.1).bfloat16()
W = (torch.randn(Hp, Hp, device=dev) / math.sqrt(H)).bfloat16()
bias = torch.zeros(Hp, device=dev)
out = torch.empty(rows, Hp, device=dev, dtype=torch.bfloat16)
dW = torch.zeros(H, H, device=dev)
for _ in range(int(os.environ.get("ITERS", 3))):
    SF.gemm_bf16(0, A, W, M=rows, N=Hp, K=Hp, bias=bias, activation=0, out=out)
    SF.gemm_bf16(1, D, W, M=rows, N=Hp, K=Hp, aux=A, activation=0, out=out)
    SF.gemm_bf16(2, D, A, M=H, N=H, K=rows, out=dW)
torch.cuda.synchronize()
print("ok")
