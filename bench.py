#!/usr/bin/env python
"""bench.py -- spatial-VAE train-step throughput on B200 (see the task contract in DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU port of the reference path

A "step" is one full train step of the hot path on one minibatch of synthetic images:
gather the shuffled batch -> encoder -> reparameterise -> rotate/translate -> per-pixel decoder ->
ELBO -> backward -> (allreduce) -> Adam.  Workload at N GPUs: BASELINE.json configs[1]
(rotated+translated MNIST 28x28, z-dim 100, p-hidden 500x2, minibatch 1024 PER GPU: weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import io
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "spatial-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CONFIGS = {
    # name: family, n (image side), Cin, C_out, Z, H, L, Hq, Lq, batch per GPU, theta_prior, ctf
    "c1": dict(family="mnist", n=28, Cin=1, C=1, Z=2, H=500, L=2, Hq=500, Lq=2, B=100, theta_prior=math.pi / 4),
    "c2": dict(family="mnist", n=28, Cin=1, C=1, Z=100, H=500, L=2, Hq=500, Lq=2, B=1024, theta_prior=math.pi / 4),
    "c3": dict(family="particles", n=40, Cin=1, C=2, Z=2, H=500, L=2, Hq=500, Lq=2, B=512, theta_prior=math.pi,
               augment=True),
    "c4": dict(family="galaxy", n=64, Cin=3, C=3, Z=20, H=1000, L=4, Hq=5000, Lq=2, B=128, theta_prior=math.pi),
    "c5": dict(family="particles", n=40, Cin=1, C=1, Z=2, H=500, L=2, Hq=500, Lq=2, B=4096, theta_prior=math.pi,
               ctf=39),
}
WORKLOAD_TEXT = {
    "c1": "train_mnist.py rotated MNIST 28x28, z-dim 2, p-hidden 500x2, minibatch 100",
    "c2": "rotated+translated MNIST 28x28, z-dim 100, p-hidden 500x2, minibatch 1024 per GPU",
    "c3": "5HDB-like EM particles 40x40, --fit-noise, --augment-rotation (device bicubic), minibatch 512 per GPU",
    "c4": "galaxy zoo 64x64x3, z-dim 20, p-hidden 1000x4, q-hidden 5000x2, minibatch 128 per GPU",
    "c5": "CODH/ACS-like EM particles 40x40 with 39x39 CTF kernels, minibatch 4096 per GPU",
}


def flops_per_image_train(c):
    """Algorithmic FLOPs of one train step per image: every Linear 2*M*K*N forward, x3 for
    fwd+dX+dW, unpadded H; recompute / padding / elementwise not counted; + CTF 2*2*P*k^2."""
    P = c["n"] * c["n"]
    I = c["Z"] + 3
    dec = P * 2 * (2 * c["H"] + (c["L"] - 1) * c["H"] ** 2 + c["H"] * c["C"]) + 2 * c["Z"] * c["H"]
    enc = 2 * (P * c["Cin"] * c["Hq"] + (c["Lq"] - 1) * c["Hq"] ** 2 + c["Hq"] * 2 * I)
    ctf = 2 * 2 * P * c.get("ctf", 0) ** 2
    return 3 * (dec + enc) + ctf


def synth_images(c, count, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    P = c["n"] * c["n"]
    if c["family"] == "mnist":       # ~80 % exact zeros, values in [0,1]
        u = torch.rand(count, P, generator=g, device=device)
        return (u > 0.8).float() * torch.rand(count, P, generator=g, device=device)
    if c["family"] == "galaxy":
        return torch.rand(count, P, 3, generator=g, device=device)
    return torch.randn(count, P, generator=g, device=device)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML during the timed region.  (Streaming
    `nvidia-smi -lms` from a side process stalled NCCL steps on multi-GPU runs; the in-process NVML
    queries of ONE device do not.)"""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.25):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.stop_flag, self.max_mhz = index, period, [], False, None
        self.query_ms = []
        if os.environ.get("BENCH_NVML_PERIOD"):
            self.period = float(os.environ["BENCH_NVML_PERIOD"])

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[0].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            what = os.environ.get("BENCH_NVML", "both")
            while not self.stop_flag:
                t0 = time.perf_counter()
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM) if what in ("both", "clock") else 0
                mask = 0
                if what in ("both", "reasons"):
                    try:
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.query_ms.append((time.perf_counter() - t0) * 1e3)
                self.samples.append((mhz, mask))
                time.sleep(self.period)
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        self.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(s[0] for s in self.samples)
        reasons = [n for n, bit in self.REASONS.items() if any(s[1] & bit for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(mhz),
                "nvml_query_ms": round(sum(self.query_ms) / len(self.query_ms), 2)}


def nvml_reasons(index):
    """(max SM MHz, set of active throttle reasons) read once through NVML, outside the timed region."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[index]) if vis and vis.split(",")[0].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        try:
            mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        return pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), {n for n, b in ClockSampler.REASONS.items() if mask & b}
    except Exception:
        return None, {"unavailable"}


def clocks_summary(sampler, probe_mhz, before, after):
    """sm_mhz = median of the on-device clock probes taken inside the timed region; reasons = union of the NVML
    samples taken during the region (single GPU) and right before / after it (always)."""
    out = {"sm_mhz": round(probe_mhz[len(probe_mhz) // 2]) if probe_mhz else None,
           "sm_mhz_min": round(probe_mhz[0]) if probe_mhz else None,
           "sm_max_mhz": (before or (None,))[0], "probes": len(probe_mhz),
           "method": "on-device clock64/globaltimer probes during the timed region"}
    reasons = set()
    for r in (before, after):
        if r:
            reasons |= r[1]
    if sampler is not None:
        s = sampler.summary()
        reasons |= set(s.get("reasons", []))
        out["nvml_sm_mhz"] = s.get("sm_mhz")
        out["nvml_samples"] = s.get("samples")
        out["method"] += " + NVML at 4 Hz during it"
    else:
        out["method"] += " + NVML right before/after it (NVML polling during NCCL steps perturbs them)"
    out["reasons"] = sorted(reasons)
    return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def cpu_reference_steps(c, batch, steps, warmup, threads):
    """The oracle port of the reference train step (eval_minibatch + backward + Adam) on host cores."""
    from oracle import svae_oracle as O
    torch.set_num_threads(threads)
    P = c["n"] * c["n"]
    dec, enc = O.init_params(P * c["Cin"], c["Z"] + 3, c["Z"], c["H"], c["L"], c["Hq"], c["Lq"], c["C"], seed=0)
    cfg = O.StepConfig(family=c["family"], theta_prior=c["theta_prior"])
    grid = O.make_grid(c["n"], c["n"])
    y = synth_images(c, batch, torch.device("cpu"), 1234)
    ctf = None
    if c.get("ctf"):
        ctf = 0.03 * torch.randn(batch, 1, c["ctf"], c["ctf"], generator=torch.Generator().manual_seed(5))
    adam = O.AdamState(lr=1e-4)
    times = []
    for s in range(warmup + steps):
        eps = torch.randn(batch, c["Z"] + 3, generator=torch.Generator().manual_seed(1000 + s))
        t0 = time.perf_counter()
        out, grads = O.step_grads(cfg, dec, enc, grid, y, eps, **({"ctf": ctf} if ctf is not None else {}))
        new = adam.update(O.flatten_params(dec, enc), grads)
        dec, enc = O.unflatten_like(dec, enc, new)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def reference_arm(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_b = min(c["B"], 64 if c["H"] <= 500 else 4)
    sec = cpu_reference_steps(c, sample_b, args.steps, max(args.warmup, 1), threads)
    val = sample_b / sec
    line = {"impl": "reference", "metric": "train images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.config], "sample": f"{sample_b} images per CPU step"},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"oracle port of eval_minibatch+backward+Adam, {sample_b} images/step, "
                                       f"{args.steps} steps"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def build_models(c, device):
    import torch.nn as nn
    import spatial_vae.models as M
    P = c["n"] * c["n"]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(c["Z"], c["H"], n_out=c["C"], num_layers=c["L"], activation=nn.Tanh)
        q = M.InferenceNetwork(P * c["Cin"], c["Z"] + 3, c["Hq"], num_layers=c["Lq"], activation=nn.Tanh)
    return p.to(device), q.to(device)


def time_gemm_kernels(c, rows, device, iters=10):
    """Device time of the three tcgen05 GEMMs of one hidden layer at the workload's row count,
    CUDA events on the launching stream (torch's current stream)."""
    import spatial_vae.functional as SF
    Hp = (c["H"] + 63) // 64 * 64
    H = c["H"]
    A = (torch.randn(rows, Hp, device=device) * 0.5).bfloat16()
    D = (torch.randn(rows, Hp, device=device) * 0.1).bfloat16()
    W = (torch.randn(Hp, Hp, device=device) / math.sqrt(H)).bfloat16()
    bias = torch.zeros(Hp, device=device)
    out = torch.empty(rows, Hp, device=device, dtype=torch.bfloat16)
    dW = torch.zeros(H, H, device=device)
    res = {}
    calls = {
        "fwd": lambda: SF.gemm_bf16(0, A, W, M=rows, N=Hp, K=Hp, bias=bias, activation=0, out=out),
        "dx": lambda: SF.gemm_bf16(1, D, W, M=rows, N=Hp, K=Hp, aux=A, activation=0, out=out),
        "dw": lambda: SF.gemm_bf16(2, D, A, M=H, N=H, K=rows, out=dW),
    }
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / iters * 1e-3
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fast", choices=["fast", "parity_tc", "parity"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the config's)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel of the step instead of replaying a CUDA graph")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.batch > 0:
        c["B"] = args.batch
    if args.impl == "reference":
        reference_arm(args, c)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import spatial_vae.functional as SF
    from spatial_vae import _lib as L
    from spatial_vae.trainer import Trainer
    from spatial_vae.driver import make_grid

    P = c["n"] * c["n"]
    B = c["B"]
    p_net, q_net = build_models(c, device)
    spec = SF.StepSpec(family=c["family"], theta_prior=c["theta_prior"], precision=args.precision,
                       chunk_images=args.chunk)
    trainer = Trainer(p_net, q_net, spec, lr=1e-4)
    grid = make_grid(c["n"], c["n"], device)
    n_data = 8 * B
    data = synth_images(c, n_data, device, 1234 + rank)
    ctf_all = None
    if c.get("ctf"):
        ctf_all = 0.03 * torch.randn(n_data, c["ctf"], c["ctf"], device=device,
                                     generator=torch.Generator(device=device).manual_seed(77 + rank))
    perm_gen = torch.Generator(device=device).manual_seed(4321)
    import numpy as np
    aug_rng = np.random.default_rng(7 + rank)

    # DataLoader(shuffle=True) semantics (reference train_mnist.py:395): ONE permutation of the dataset per epoch,
    # consecutive minibatches are consecutive slices of it (8 steps per epoch here)
    epoch_state = {"perm": None, "pos": 0}

    def next_indices():
        if epoch_state["perm"] is None or epoch_state["pos"] + B > n_data:
            epoch_state["perm"] = torch.randperm(n_data, generator=perm_gen, device=device)
            epoch_state["pos"] = 0
        lo = epoch_state["pos"]
        epoch_state["pos"] = lo + B
        return epoch_state["perm"][lo:lo + B]

    def device_step(i):
        idx = next_indices()
        y = SF.gather_rows(data, idx)
        ctf = SF.gather_rows(ctf_all, idx) if ctf_all is not None else None
        y_enc = theta_offset = None
        if c.get("augment"):      # --augment-rotation: encoder sees a randomly rotated copy, decoder theta gets the offset
            offs = aug_rng.uniform(0, 2 * math.pi, size=B)
            y_enc = SF.rotate_bicubic(y, c["n"], c["n"], offs * (360 / 2 / math.pi))
            theta_offset = torch.from_numpy(offs).float().to(device, non_blocking=True)
        return step_fn(grid, y, global_batch=B * world, ctf=ctf, y_enc=y_enc, theta_offset=theta_offset)

    # CUDA-graph replay on one GPU.  With several ranks the step is enqueued kernel by kernel: capturing the NCCL
    # allreduce worked at 8 ranks on C4/C5, but an 8-rank C2 run hung in the same session and there was no GPU budget
    # left to find out why, so the multi-GPU default stays on the path that was measured at 2/4/8 GPUs.
    use_graph = (not args.no_graph) and (world == 1 or os.environ.get("BENCH_GRAPH_MULTI") == "1")
    step_fn = trainer.step_graphed if use_graph else trainer.step
    for i in range(args.warmup):
        device_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # NVML polling from a side thread measurably slows NCCL steps (+13 % at 2 GPUs even at 1 Hz), so with more
    # than one rank the clock under load comes from on-device probes and NVML is read right before / after.
    sampler = ClockSampler(local) if (rank == 0 and world == 1 and not os.environ.get("BENCH_NO_SAMPLER")) else None
    probe_every = max(1, args.steps // 8)
    probes = torch.zeros(args.steps // probe_every + 2, dtype=torch.float32, device=device)
    reasons_before = nvml_reasons(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.1)           # first NVML sample lands before the timed region starts
    # library launches of ONE step (counted on an eager step; a graph replay re-issues the same kernels)
    lc0 = L.lib.svae_launch_count()
    trainer.step(grid, SF.gather_rows(data, torch.arange(B, device=device)), global_batch=B * world,
                 ctf=SF.gather_rows(ctf_all, torch.arange(B, device=device)) if ctf_all is not None else None)
    launches_per_step = L.lib.svae_launch_count() - lc0        # 2 gathers + the step's kernels + Adam
    torch.cuda.synchronize()
    launches0 = L.lib.svae_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n_probe = 0
    for i in range(args.steps):
        res = device_step(i)
        if i % probe_every == probe_every // 2:      # ~8 probes of 20 us spread over the timed region
            L.check(L.lib.svae_sm_clock_probe(probes[n_probe:].data_ptr(), torch.cuda.current_stream().cuda_stream),
                    "svae_sm_clock_probe")
            n_probe += 1
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = L.lib.svae_launch_count() - launches0
    if use_graph:
        launches = launches_per_step * args.steps          # kernels replayed from the captured graph + the gathers
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    if sampler:
        sampler.stop()
    reasons_after = nvml_reasons(local) if rank == 0 else None
    probe_mhz = sorted(float(v) for v in probes[:n_probe].cpu())
    last = [float(v) for v in res.cpu()]
    # CPU time to ENQUEUE one step (short burst on an idle queue, so the launch queue never fills)
    torch.cuda.synchronize()
    t_host0 = time.perf_counter()
    for i in range(8):
        device_step(i)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / 8
    torch.cuda.synchronize()

    # ---- end to end through the public API with HOST buffers ---------------------------------------
    host_data = synth_images(c, 2 * B, torch.device("cpu"), 99 + rank).pin_memory()
    host_ctf = None
    if c.get("ctf"):
        host_ctf = (0.03 * torch.randn(2 * B, c["ctf"], c["ctf"])).pin_memory()
    out_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        lo = (i % 2) * B
        y = host_data[lo:lo + B].to(device, non_blocking=True)
        ctf = host_ctf[lo:lo + B].to(device, non_blocking=True) if host_ctf is not None else None
        y_enc = theta_offset = None
        if c.get("augment"):
            offs = aug_rng.uniform(0, 2 * math.pi, size=B)
            y_enc = SF.rotate_bicubic(y, c["n"], c["n"], offs * (360 / 2 / math.pi))
            theta_offset = torch.from_numpy(offs).float().to(device, non_blocking=True)
        r = step_fn(grid, y, global_batch=B * world, ctf=ctf, y_enc=y_enc, theta_offset=theta_offset)
        out_host.copy_(r, non_blocking=False)      # device -> host read of the step's result (syncs)
        return out_host

    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    k2 = max(5, args.steps // 2)
    for i in range(k2):
        e2e_step(i)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = B * world * k2 / float(t_e2e)
    h2d = B * P * c["Cin"] * 4 + (B * c["ctf"] ** 2 * 4 if c.get("ctf") else 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    burst, sustained, hbm, src = peaks()
    fl_img = flops_per_image_train(c)
    line = {
        "metric": "train images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "fast" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.config], "images_per_gpu": B, "global_batch": B * world,
                   "precision": args.precision + (" (bf16 tcgen05 hidden GEMMs, fp32 accumulate; everything else fp32)"
                                                  if args.precision == "fast" else " (fp32 FFMA)"),
                   "parallelism": f"dp{world}",
                   "launch": "CUDA graph replay of the captured step" if use_graph else "eager",
                   "l2": "no explicit flush: each step streams >2 GB of activations (>> 126 MB L2) and "
                         "gathers a fresh shuffled batch (one permutation per 8-step epoch, as DataLoader(shuffle=True))"},
        "pixel_evals_per_s": value * P,
        "step_tflops_algorithmic": value * fl_img / 1e12,
        "step_frac_of_sustained_bf16": value * fl_img / 1e12 / (sustained * world),
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": host_ms,
        "last_step": {"elbo": last[0], "logp": last[1], "kl": last[2]},
        "clocks": clocks_summary(sampler, probe_mhz, reasons_before, reasons_after),
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12},
    }
    if args.precision == "fast" and c["L"] >= 2:
        rows = B * P
        kt = time_gemm_kernels(c, rows, device)
        alg = 2.0 * rows * c["H"] * c["H"]
        dom = max(kt, key=kt.get)
        # DRAM bytes per launch (read + write) of the three GEMMs from the committed `ncu --set full` capture
        # (profiles/r01_v3_summary.md), which was taken at exactly this row count and width; null for other shapes
        traffic = None
        if rows == 1024 * 784 and c["H"] == 500:
            traffic = {"fwd": 0.822635e9 + 0.777726e9, "dx": 1.644998e9 + 0.791843e9, "dw": 1.645260e9 + 0.004575e9}[dom]
        line["roofline"] = {"bound": "tensor", "kernel": f"tc_gemm_kernel<{dom}>", "achieved": alg / kt[dom] / 1e12,
                            "peak": burst, "unit": "TFLOP/s", "frac": alg / kt[dom] / 1e12 / burst, "traffic": traffic,
                            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                              "profiles/r01_v3_summary.md" if traffic else None,
                            "peak_source": f"MEASURED_PEAKS.json bf16_tflops (burst), {src}",
                            "all_kernels_tflops": {k: alg / v / 1e12 for k, v in kt.items()},
                            "all_kernels_ms": {k: v * 1e3 for k, v in kt.items()}}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sb = min(B, 64 if c["H"] <= 500 else 4)
        sec = cpu_reference_steps(c, sb, 5, 2, threads)
        line["cpu_baseline"] = {"value": sb / sec, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port of eval_minibatch+backward+Adam, {sb} images/step, 5 steps"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
