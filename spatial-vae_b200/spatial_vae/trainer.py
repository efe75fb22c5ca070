"""Train-step driver: what the reference's train_epoch body does per minibatch
(train_mnist.py:138-150: eval_minibatch, loss.backward(), optim.step(), optim.zero_grad()),
as one fused pass over flat parameter / gradient / Adam buffers, data-parallel over the
ranks of torch.distributed with ONE gradient allreduce per step (NCCL over NVLink on a B200 box,
gloo in the CPU tests of the host logic).

The nn.Parameters of p_net / q_net are re-pointed at views of one flat fp32 buffer, in the
optimiser order of the reference (p_net.parameters() then q_net.parameters(), train_mnist.py:387),
so state_dict()/torch.save keep working and Adam is a single kernel over the flat buffer.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import functional as SF


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, near-even split of n items over `world` ranks (ragged tail allowed, a rank may
    get 0 items): the data-parallel partition of one minibatch."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class FlatParams:
    """Flat fp32 storage for a list of parameters, plus same-shaped flat grad / Adam moments."""

    def __init__(self, params):
        params = list(params)
        self.shapes = [p.shape for p in params]
        self.numels = [p.numel() for p in params]
        dev = params[0].device
        # pad every tensor to a multiple of 4 floats so each view is 16-byte aligned
        self.offsets, off = [], 0
        for n in self.numels:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4
        self.total = off
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        # the gradient buffer carries 4 extra floats: the per-rank loss sums (logp, kl, elbo, unused) ride in the
        # same allreduce as the gradient
        self.grad_ext = torch.zeros(off + 4, dtype=torch.float32, device=dev)
        self.grad = self.grad_ext[:off]
        self.m = torch.zeros_like(self.data)
        self.v = torch.zeros_like(self.data)
        for p, o, n in zip(params, self.offsets, self.numels):
            self.data[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.data[o:o + n].view(p.shape)
        self.params = params

    def views(self, flat):
        return [flat[o:o + n].view(s) for o, n, s in zip(self.offsets, self.numels, self.shapes)]


class Trainer:
    """One object per process (= per GPU).  `step(y, ...)` runs the whole train step on this rank's
    slice of the global minibatch and leaves (elbo, logp, kl) batch means on the device."""

    def __init__(self, p_net, q_net, spec: SF.StepSpec, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None, seed: Optional[int] = None):
        if bool(getattr(p_net, "resid", False)) != bool(getattr(q_net, "resid", False)):
            raise NotImplementedError("one resid flag for both networks, as the reference's --resid")
        if bool(getattr(p_net, "resid", False)) != bool(spec.resid):
            spec = SF.StepSpec(**{**spec.__dict__, "resid": bool(getattr(p_net, "resid", False))})
        self.p_net, self.q_net, self.spec = p_net, q_net, spec
        self.lr, self.betas, self.adam_eps = lr, betas, eps
        self.flat = FlatParams(list(p_net.parameters()) + list(q_net.parameters()))
        self.t = 0                                  # host mirror of the device-resident Adam step counter
        dev = self.flat.data.device
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bc_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self._graphs = {}
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        if self.world > 1:
            # data-parallel replicas start from rank 0's weights whatever each process initialised (the reference
            # never seeds, train_mnist.py has no manual_seed); Adam moments and the step counter start at zero
            dist.broadcast(self.flat.data, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                           group=process_group)
        # eps: drawn in the kernel (Philox keyed on seed, step, GLOBAL image index), so an N-rank run draws the
        # same numbers as a single-rank run of the same minibatch.  The seed is shared by all ranks.
        if seed is None:
            seed_t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
            if self.world > 1:
                seed_t = seed_t.to(dev)
                dist.broadcast(seed_t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)
            seed = int(seed_t.item())
        self.seed = seed
        # Optional two-bucket gradient exchange on a GPU (SVAE_SPLIT_ALLREDUCE=1): the decoder's gradients are final
        # before the encoder backward starts (the library records an event there), so their allreduce runs on a second
        # stream under the encoder backward; the encoder's gradients and the loss sums follow on the main stream.
        # Measured at C2 inside the captured graph: 1.706 -> 1.686 ms at 2 GPUs, 1.728 -> 1.743 ms at 8 (one run; the
        # second collective's latency costs what the overlap gains), hence off by default: ONE allreduce per step.
        self.split_allreduce = (self.world > 1 and dev.type == "cuda" and
                                os.environ.get("SVAE_SPLIT_ALLREDUCE", "0") == "1")
        self._comm = self._dec_ready = None
        if self.split_allreduce:
            self._comm = torch.cuda.Stream(device=dev)
            self._dec_ready = torch.cuda.Event()
            self._dec_ready.record()            # creates the CUDA handle
        self._bind()

    def _bind(self):
        dec = SF.decoder_tensors_of(self.p_net)
        enc = SF.encoder_pairs_of(self.q_net)
        self.dec, self.enc = dec, enc
        n_dec = len(dec.flat())
        gv = self.flat.views(self.flat.grad)
        self.gdec = SF.DecoderTensors.from_flat(gv[:n_dec], *dec.layout())
        self.genc = [(gv[n_dec + i], gv[n_dec + i + 1]) for i in range(0, len(gv) - n_dec, 2)]
        self.n_dec_flat = self.flat.offsets[n_dec]          # the decoder's share of the flat buffers comes first

    # -- one train step ---------------------------------------------------------------------------
    def step(self, x_coord: torch.Tensor, y_local: torch.Tensor, *, global_batch: Optional[int] = None,
             eps: Optional[torch.Tensor] = None, y_enc=None, theta_offset=None, ctf=None, mask=None,
             z_scale: Optional[float] = None, image_offset: Optional[int] = None) -> torch.Tensor:
        """y_local: this rank's images.  global_batch: images over all ranks (default: local size x world
        when every rank holds the same count).  eps None: drawn in the kernel (see __init__); image_offset is the
        global minibatch index of y_local[0] (default: rank * local size).  Returns a device tensor [elbo, logp, kl]
        (global batch means); nothing is synchronised with the host."""
        B_global = global_batch if global_batch is not None else y_local.shape[0] * self.world
        out = self._step_body(x_coord, y_local, B_global, eps, y_enc, theta_offset, ctf, mask, z_scale, image_offset)
        self.t += 1
        return out

    # -- the same step replayed from a CUDA graph ------------------------------------------------------
    def step_graphed(self, x_coord: torch.Tensor, y_local: torch.Tensor, *, global_batch: Optional[int] = None,
                     ctf=None, mask=None, y_enc=None, theta_offset=None, z_scale: Optional[float] = None,
                     image_offset: Optional[int] = None) -> torch.Tensor:
        """`step` with the ~40 kernel launches (+ the NCCL allreduce) of one train step captured once per input
        shape into a CUDA graph and replayed: the inputs are copied into static buffers, eps is drawn inside the
        graph, Adam's step counter and bias corrections live in device memory.  The first call for a
        shape runs one ordinary eager step (lazy initialisation must not happen under capture) and captures on the
        second.  Returns the same [elbo, logp, kl] tensor (a static buffer: copy it if you keep it)."""
        B_local = y_local.shape[0]
        B_global = global_batch if global_batch is not None else B_local * self.world
        key = (tuple(y_local.shape), B_global, None if ctf is None else tuple(ctf.shape), mask is not None,
               y_enc is not None, theta_offset is not None, z_scale, image_offset)
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "warm"
            return self.step(x_coord, y_local, global_batch=global_batch, ctf=ctf, mask=mask, y_enc=y_enc,
                             theta_offset=theta_offset, z_scale=z_scale, image_offset=image_offset)
        if entry == "warm":
            st = {"y": y_local.clone(), "ctf": ctf.clone() if ctf is not None else None,
                  "y_enc": y_enc.clone() if y_enc is not None else None,
                  "toff": theta_offset.clone() if theta_offset is not None else None}
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["out"] = self._step_body(x_coord, st["y"], B_global, None, st["y_enc"], st["toff"], st["ctf"], mask,
                                            z_scale, image_offset)
            st["graph"] = graph
            # the captured kernels hold raw pointers into the library workspace: keep that buffer alive even if a
            # later, larger request makes spatial_vae.functional.workspace() allocate a new one
            st["keepalive"] = list(SF._workspaces.values())
            self._graphs[key] = entry = st
            # the capture itself did not execute anything
        st = entry
        st["y"].copy_(y_local, non_blocking=True)
        if ctf is not None:
            st["ctf"].copy_(ctf, non_blocking=True)
        if y_enc is not None:
            st["y_enc"].copy_(y_enc, non_blocking=True)
        if theta_offset is not None:
            st["toff"].copy_(theta_offset, non_blocking=True)
        self.t += 1
        st["graph"].replay()
        return st["out"]

    def _step_body(self, x_coord, y_local, B_global, eps, y_enc, theta_offset, ctf, mask, z_scale, image_offset=None):
        """Everything of one train step that is enqueued on the stream (shared by step_graphed's capture)."""
        B_local = y_local.shape[0]
        spec = self.spec
        if z_scale is not None and z_scale != spec.z_scale:
            spec = SF.StepSpec(**{**spec.__dict__, "z_scale": z_scale})
        rng = None
        if eps is None:
            off = image_offset if image_offset is not None else self.rank * B_local
            rng = (self.seed, self.t_dev, off)          # t_dev: Adam's device-resident step counter (steps done so far)
        extra = {"decoder_grads_event": self._dec_ready} if self.split_allreduce else {}
        # the library writes this rank's three loss sums straight into the tail of the gradient buffer
        tail = self.flat.grad_ext[self.flat.total:]
        if B_local > 0:
            extra["stats_sum"] = tail
        SF.run_step(spec, self.dec, self.enc, x_coord, y_local, eps, y_enc=y_enc, theta_offset=theta_offset, ctf=ctf,
                    mask=mask, grad_dec=self.gdec, grad_enc=self.genc, grad_scale=1.0 / max(B_global, 1), rng=rng,
                    **extra)
        if self.split_allreduce:
            # first bucket: the decoder's gradients, on the second stream, as soon as the library's event fires (a
            # rank with an empty slice enqueued nothing: it joins the exchange after whatever the main stream holds)
            main = torch.cuda.current_stream()
            if B_local > 0:
                self._comm.wait_event(self._dec_ready)
            else:
                self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(self.flat.grad_ext[:self.n_dec_flat], op=dist.ReduceOp.SUM, group=self.pg)
        if B_local == 0:
            tail.zero_()
        if self.split_allreduce:
            # second bucket: the encoder's gradients and the three loss sums
            dist.all_reduce(self.flat.grad_ext[self.n_dec_flat:], op=dist.ReduceOp.SUM, group=self.pg)
            torch.cuda.current_stream().wait_stream(self._comm)
        elif self.world > 1:
            # the one exchange step of the path: gradient sum and the three loss sums in ONE allreduce
            dist.all_reduce(self.flat.grad_ext, op=dist.ReduceOp.SUM, group=self.pg)
        means = tail[:3] / max(B_global, 1)
        out = torch.stack([means[2], means[0], means[1]])
        SF.adam_step_graph(self.flat.data, self.flat.grad, self.flat.m, self.flat.v, self.lr, self.t_dev, self.bc_dev,
                           self.betas, self.adam_eps, zero_grad=True)
        return out

    @torch.no_grad()
    def evaluate(self, x_coord, y_local, *, global_batch=None, eps=None, ctf=None, mask=None, want_y_hat=False,
                 z_scale: Optional[float] = None):
        B_local = y_local.shape[0]
        B_global = global_batch if global_batch is not None else B_local * self.world
        spec = self.spec
        if z_scale is not None and z_scale != spec.z_scale:
            spec = SF.StepSpec(**{**spec.__dict__, "z_scale": z_scale})
        I = self.enc[-1][0].shape[0] // 2
        if eps is None:
            eps = torch.empty(B_local, I, dtype=torch.float32, device=y_local.device).normal_()
        stats, y_hat, _ = SF.run_step(spec, self.dec, self.enc, x_coord, y_local, eps, ctf=ctf, mask=mask,
                                      want_y_hat=want_y_hat)
        sums = stats.sum(0) if B_local > 0 else stats.new_zeros(3)
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.pg)
        means = sums / max(B_global, 1)
        return torch.stack([means[2], means[0], means[1]]), y_hat
