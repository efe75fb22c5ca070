"""Run one tensor-core building block once at the C2 row count (for ncu): python scripts/run_one.py dw_top|dw_top_nostore|dx_moments|fwd|dx|dw"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spatial-vae_b200")]
import torch
import spatial_vae.functional as SF
which = sys.argv[1]
rows, H, P = 1024 * 784, 500, 784
Hp = 512
dev = torch.device("cuda")
A = (torch.randn(rows, Hp, device=dev) * 0.5).bfloat16()
D = (torch.randn(rows, Hp, device=dev) * 0.1).bfloat16()
W = (torch.randn(Hp, Hp, device=dev) / math.sqrt(H)).bfloat16()
g_o = torch.randn(rows, 1, device=dev) * 0.1
out_w = torch.randn(1, H, device=dev) / math.sqrt(H)
B = rows // P
grid = torch.rand(P, 2, device=dev) * 2 - 1
img = torch.rand(B, 4, device=dev); cw = torch.randn(H, 2, device=dev); hz = torch.randn(B, Hp, device=dev)
bias = torch.zeros(Hp, device=dev); out = torch.empty(rows, Hp, device=dev, dtype=torch.bfloat16); dW = torch.zeros(H, H, device=dev)
fn = {"dw_top": lambda: SF.gemm_dw_top(A, D, g_o, out_w, H=H),
      "dw_top_nostore": lambda: SF.gemm_dw_top(A, D, g_o, out_w, H=H, want_delta=False),
      "dx_moments": lambda: SF.gemm_dx_moments(D, W, H=H, grid=grid, img=img, coord_w=cw, hz=hz, P=P),
      "fwd": lambda: SF.gemm_bf16(0, A, W, M=rows, N=Hp, K=Hp, bias=bias, activation=0, out=out),
      "dx": lambda: SF.gemm_bf16(1, D, W, M=rows, N=Hp, K=Hp, aux=A, activation=0, out=out),
      "dw": lambda: SF.gemm_bf16(2, D, A, M=H, N=H, K=rows, out=dW)}[which]
for _ in range(3):
    fn()
torch.cuda.synchronize()
