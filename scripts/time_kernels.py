"""Device times (CUDA events, torch's current stream) of the tensor-core kernels at a config's row count.
   python scripts/time_kernels.py [rows] [H]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spatial-vae_b200")]
import torch
import spatial_vae.functional as SF

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024 * 784
H = int(sys.argv[2]) if len(sys.argv) > 2 else 500
P = 784
Hp = (H + 63) // 64 * 64
dev = torch.device("cuda")
A = (torch.randn(rows, Hp, device=dev) * 0.5).bfloat16()
D = (torch.randn(rows, Hp, device=dev) * 0.1).bfloat16()
W = (torch.randn(Hp, Hp, device=dev) / math.sqrt(H)).bfloat16()
bias = torch.zeros(Hp, device=dev)
out = torch.empty(rows, Hp, device=dev, dtype=torch.bfloat16)
dW = torch.zeros(H, H, device=dev)
B = rows // P
grid = torch.rand(P, 2, device=dev) * 2 - 1
img = torch.rand(B, 4, device=dev)
cw = torch.randn(H, 2, device=dev)
hz = torch.randn(B, Hp, device=dev)
calls = {
    "fwd": lambda: SF.gemm_bf16(0, A, W, M=rows, N=Hp, K=Hp, bias=bias, activation=0, out=out),
    "dx": lambda: SF.gemm_bf16(1, D, W, M=rows, N=Hp, K=Hp, aux=A, activation=0, out=out),
    "dw": lambda: SF.gemm_bf16(2, D, A, M=H, N=H, K=rows, out=dW),
    "dx_moments": lambda: SF.gemm_dx_moments(D[:B * P], W, H=H, grid=grid, img=img, coord_w=cw, hz=hz, P=P),
}
g_o = torch.randn(rows, 1, device=dev) * 0.1
out_w = torch.randn(1, H, device=dev) / math.sqrt(H)
calls["dw_top"] = lambda: SF.gemm_dw_top(A, D, g_o, out_w, H=H)
calls["dw_top_nostore"] = lambda: SF.gemm_dw_top(A, D, g_o, out_w, H=H, want_delta=False)
for extra in getattr(SF, "EXTRA_TIMED_KERNELS", []):
    calls.update(extra(locals()))
alg = 2.0 * rows * H * H
for name, fn in calls.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:12s} {ms*1e3:8.1f} us   {alg/ms/1e9:7.1f} TFLOP/s algorithmic")
