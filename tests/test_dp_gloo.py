"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2.

The kernels cannot run here, so the two library calls of Trainer.step (run_step, adam_step) are
replaced by test doubles built on the CPU oracle; everything else is the product code: flat
parameter/gradient buffers behind the nn.Parameters, the contiguous ragged sharding of a minibatch,
grad_scale = 1/B_global, the one gradient allreduce and the loss-sum allreduce, replicated Adam.
The result must equal the single-process oracle trajectory on the full minibatch.
"""
import contextlib
import io
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import svae_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _install_doubles(SF):
    def fake_run_step(spec, dec, enc, grid, y, eps, *, y_enc=None, theta_offset=None, ctf=None, mask=None,
                      grad_dec=None, grad_enc=None, grad_scale=None, want_y_hat=False, want_latent=False, rng=None,
                      stats_sum=None):
        B = y.shape[0]
        if B == 0:
            return torch.zeros(0, 3), None, None
        if eps is None:      # stand-in for the in-kernel draw: a function of (seed, step, GLOBAL image index) only
            seed, step_t, off = rng
            I = enc[-1][0].shape[0] // 2
            eps = torch.stack([torch.randn(I, generator=torch.Generator().manual_seed(
                (seed + 1000003 * int(step_t) + 7919 * (off + b)) % (2 ** 31))) for b in range(B)])
        cfg = O.StepConfig(family=spec.family, rotate=spec.rotate, translate=spec.translate, dx_scale=spec.dx_scale,
                           theta_prior=spec.theta_prior, z_scale=spec.z_scale, softplus=spec.softplus,
                           resid=spec.resid)
        d = {"coord_w": dec.coord_w.detach(), "coord_b": dec.coord_b.detach(),
             "latent_w": dec.latent_w.detach() if dec.latent_w is not None else None,
             "hidden": [(w.detach(), b.detach()) for w, b in dec.hidden], "out_w": dec.out_w.detach(),
             "out_b": dec.out_b.detach()}
        if dec.bilinear_w is not None:
            d["bilinear_w"] = dec.bilinear_w.detach()
        if spec.resid:
            d["resid"] = True
        e = [(w.detach(), b.detach()) for w, b in enc]
        out, grads = O.step_grads(cfg, d, e, grid, y, eps)
        if grad_dec is not None:
            scale = (grad_scale if grad_scale is not None else 1.0 / B) * B    # oracle grads are of the batch MEAN
            targets = grad_dec.flat() + [t for pair in grad_enc for t in pair]
            for t, g in zip(targets, grads):
                t.add_(g * scale)
        stats = torch.stack([out["logp_i"], out["kl_i"], out["logp_i"] - out["kl_i"]], 1)
        if stats_sum is not None:        # what the library's finalize kernel writes
            stats_sum[:3] = stats.sum(0)
            stats_sum[3] = 0
        return stats, None, None

    def fake_adam(param, grad, m, v, lr, t_dev, bc_dev, betas=(0.9, 0.999), eps=1e-8, zero_grad=True):
        t_dev += 1                       # the device-resident step counter of svae_adam_tick
        t = int(t_dev)
        m.mul_(betas[0]).add_(grad, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(grad, grad, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        param.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))
        if zero_grad:
            grad.zero_()

    SF.run_step = fake_run_step
    SF.adam_step_graph = fake_adam


def _worker(rank, world, port, batches, tmp, backend="doubles"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spatial_vae.functional as SF
        import spatial_vae.models as M
        from spatial_vae.trainer import Trainer, shard_bounds
        if backend == "doubles":
            _install_doubles(SF)
        else:       # the library's own kernel sources, compiled for the host (tests/simt_emu)
            from tests import emu_backend
            emu_backend.install(pytest.MonkeyPatch())
        torch.manual_seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            p = M.SpatialGenerator(3, 16, n_out=1, num_layers=2)
            q = M.InferenceNetwork(36, 6, 12, num_layers=2)
        spec = SF.StepSpec(family="mnist", theta_prior=0.7, precision="parity")
        tr = Trainer(p, q, spec, lr=1e-3)
        assert tr.world == world and tr.rank == rank
        grid = O.make_grid(6, 6)
        losses = []
        for y, eps in batches:
            lo, hi = shard_bounds(y.shape[0], world, rank)
            res = tr.step(grid, y[lo:hi], global_batch=y.shape[0], eps=eps[lo:hi])
            losses.append(res.clone())
        # parameters are views of the flat buffer and identical on every rank
        assert p.coord_linear.weight.data_ptr() == tr.flat.data.data_ptr()
        flat = tr.flat.data.clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered)
        if rank == 0:
            torch.save({"state_p": p.state_dict(), "state_q": q.state_dict(), "losses": torch.stack(losses)}, tmp)
    finally:
        dist.destroy_process_group()


def _batches():
    g = torch.Generator().manual_seed(5)
    out = []
    for B in (5, 4, 1):        # ragged split 3+2, even split, and a rank with an empty slice
        y = (torch.rand(B, 36, generator=g) > 0.7).float() * torch.rand(B, 36, generator=g)
        out.append((y, torch.randn(B, 6, generator=g)))
    return out


@pytest.mark.parametrize("backend", ["doubles", "simt_emu"])
def test_two_rank_trainer_matches_single_process_oracle(tmp_path, backend):
    """backend "doubles": library calls replaced by oracle-backed doubles (pure host logic); "simt_emu": the real
    svae_step / Adam entry points running the kernel sources on the CPU, so the per-rank grad_scale, the ragged and
    empty shards and the gradient allreduce are checked through the C ABI itself."""
    batches = _batches()
    tmp = str(tmp_path / "dp.pt")
    if backend == "simt_emu":
        from tests.simt_emu.build import build
        build()                                  # compile once here, not concurrently in both ranks
    mp.spawn(_worker, args=(2, _free_port(), batches, tmp, backend), nprocs=2, join=True)
    got = torch.load(tmp)

    # single-process oracle trajectory on the full minibatches, from the same initial parameters
    import spatial_vae.models as M
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(3, 16, n_out=1, num_layers=2)
        q = M.InferenceNetwork(36, 6, 12, num_layers=2)
    dec = O.decoder_params_from_state({k: v.clone() for k, v in p.state_dict().items()})
    enc = O.encoder_params_from_state({k: v.clone() for k, v in q.state_dict().items()})
    cfg = O.StepConfig(family="mnist", theta_prior=0.7)
    dec2, enc2, elbos = O.train_steps(cfg, dec, enc, O.make_grid(6, 6), [b[0] for b in batches],
                                      [b[1] for b in batches], lr=1e-3)
    np.testing.assert_allclose(got["losses"][:, 0].numpy(), elbos, rtol=1e-5)
    ref = O.flatten_params(dec2, enc2)
    mine = list(got["state_p"].values()) + list(got["state_q"].values())
    for a, b in zip(mine, ref):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-6)


def test_shard_bounds_cover_and_are_contiguous():
    from spatial_vae.trainer import shard_bounds
    for n in (0, 1, 5, 8, 1023, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= max(1, (n + w - 1) // w)


def test_trainer_host_logic_with_every_option(monkeypatch):
    """Single process, library calls replaced by the oracle doubles: the flat parameter / gradient views must line
    up with the parameter order of a resid + expand_coords + bilinear + softplus network pair
    (coord W,b; latent W; bilinear W; layers.N.linear W,b; out W,b; then the encoder), so that 5 trainer steps
    equal 5 oracle steps."""
    import spatial_vae.functional as SF
    import spatial_vae.models as M
    from spatial_vae.trainer import Trainer
    monkeypatch.setattr(SF, "run_step", SF.run_step)
    monkeypatch.setattr(SF, "adam_step_graph", SF.adam_step_graph)
    _install_doubles(SF)
    torch.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(3, 16, n_out=2, num_layers=3, resid=True, expand_coords=True, bilinear=True,
                               softplus=True)
        q = M.InferenceNetwork(36, 6, 12, num_layers=3, resid=True)
    dec0 = O.decoder_params_from_state({k: v.detach().clone() for k, v in p.state_dict().items()})
    enc0 = O.encoder_params_from_state({k: v.detach().clone() for k, v in q.state_dict().items()})
    assert dec0["coord_w"].shape[1] == 5 and dec0["bilinear_w"] is not None and dec0.get("resid")
    spec = SF.StepSpec(family="particles", theta_prior=0.9, precision="parity", softplus=True, z_scale=0.5)
    tr = Trainer(p, q, spec, lr=1e-3)
    assert tr.spec.resid                               # picked up from the networks
    grid = O.make_grid(6, 6)
    g = torch.Generator().manual_seed(2)
    ys = [torch.randn(4, 36, generator=g) for _ in range(5)]
    eps = [torch.randn(4, 6, generator=g) for _ in range(5)]
    for y, e in zip(ys, eps):
        tr.step(grid, y, eps=e)
    cfg = O.StepConfig(family="particles", theta_prior=0.9, softplus=True, resid=True, z_scale=0.5)
    dec_o, enc_o, _ = O.train_steps(cfg, dec0, enc0, grid, ys, eps, lr=1e-3)
    for t, r in zip(list(p.parameters()) + list(q.parameters()), O.flatten_params(dec_o, enc_o)):
        np.testing.assert_allclose(t.detach().numpy(), r.numpy(), rtol=1e-5, atol=1e-6)


def _epoch_worker(rank, world, port, tmp, family):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spatial_vae.functional as SF
        import spatial_vae.models as M
        from spatial_vae import driver as D
        from spatial_vae.trainer import Trainer
        from tests import emu_backend
        emu_backend.install_all(pytest.MonkeyPatch())
        seen = []
        real_gather = SF.gather_rows

        def recording_gather(src, index, out=None):
            if src.shape[1] == 36:                      # the image tensor (not a CTF stack)
                seen.append(index.clone())
            return real_gather(src, index, out)

        SF.gather_rows = recording_gather
        torch.manual_seed(7)                            # identical initial weights on both ranks
        with contextlib.redirect_stdout(io.StringIO()):
            p = M.SpatialGenerator(2, 16, n_out=1, num_layers=2)
            q = M.InferenceNetwork(36, 5, 12, num_layers=2)
        torch.manual_seed(100 + rank)                   # but each rank draws its own eps
        spec = SF.StepSpec(family=family, theta_prior=0.7, precision="parity")
        tr = Trainer(p, q, spec, lr=1e-3)
        g = torch.Generator().manual_seed(3)
        data = (torch.rand(21, 36, generator=g) > 0.7).float() * torch.rand(21, 36, generator=g)
        shuffle = torch.Generator().manual_seed(11)     # the same CPU permutation on every rank
        extra = {}
        if family == "particles":                       # per-image CTF kernels + rotation augmentation + z_scale
            extra = dict(ctf=0.05 * torch.randn(21, 5, 5, generator=g), z_scale=0.5,
                         augment=lambda y: D._augment(y, True, True, 1))
        res = D.run_epoch(tr, O.make_grid(6, 6), data, train=True, minibatch_size=10, generator=shuffle,
                          progress=False, **extra)
        extra.pop("augment", None)
        val = D.run_epoch(tr, O.make_grid(6, 6), data, train=False, minibatch_size=10, progress=False, **extra)
        flat = tr.flat.data.clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered), "replicated parameters diverged"
        stats = torch.tensor(list(res) + list(val), dtype=torch.float64)
        both = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(both, stats)
        assert all(torch.equal(both[0], t) for t in both), "ranks report different epoch means"
        torch.save({"seen": torch.cat(seen), "stats": stats}, f"{tmp}.{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("family", ["mnist", "particles"])
def test_two_rank_run_epoch_on_the_simt_emulation(tmp_path, family):
    """driver.run_epoch under 2 ranks (gloo) through the real C ABI on tests/simt_emu: every image is visited exactly
    once per epoch across the ranks (21 images, minibatch 10: the last minibatch has ONE image, so rank 1 gets an
    empty slice), the replicated parameters stay bit-identical and both ranks report the same epoch means."""
    from tests.simt_emu.build import build
    build()
    tmp = str(tmp_path / "epoch.pt")
    mp.spawn(_epoch_worker, args=(2, _free_port(), tmp, family), nprocs=2, join=True)
    r0, r1 = torch.load(tmp + ".0"), torch.load(tmp + ".1")
    seen = torch.cat([r0["seen"], r1["seen"]])
    counts = torch.bincount(seen, minlength=21)
    assert torch.equal(counts, torch.full((21,), 2)), counts   # once in the training pass, once in validation
    assert torch.isfinite(r0["stats"]).all()


def _bcast_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spatial_vae.functional as SF
        import spatial_vae.models as M
        from spatial_vae.trainer import Trainer, shard_bounds
        from tests import emu_backend
        emu_backend.install(pytest.MonkeyPatch())
        torch.manual_seed(50 + rank)                     # every rank initialises DIFFERENT weights (no --seed)
        with contextlib.redirect_stdout(io.StringIO()):
            p = M.SpatialGenerator(3, 16, n_out=1, num_layers=2)
            q = M.InferenceNetwork(36, 6, 12, num_layers=2)
        before = torch.cat([t.detach().reshape(-1).clone() for t in list(p.parameters()) + list(q.parameters())])
        tr = Trainer(p, q, SF.StepSpec(family="mnist", theta_prior=0.7, precision="parity"), lr=1e-3)
        flat = tr.flat.data.clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered), "replicas differ after construction"
        seeds = [None] * world
        dist.all_gather_object(seeds, tr.seed)
        assert len(set(seeds)) == 1, "ranks must share the eps seed"
        # in-kernel eps: two steps on a ragged split; the result must not depend on the split
        g = torch.Generator().manual_seed(9)
        grid = O.make_grid(6, 6)
        out = []
        for B in (5, 4):
            y = (torch.rand(B, 36, generator=g) > 0.7).float() * torch.rand(B, 36, generator=g)
            lo, hi = shard_bounds(B, world, rank)
            out.append(tr.step(grid, y[lo:hi], global_batch=B, image_offset=lo).clone())
        if rank == 0:
            torch.save({"before_rank0": before, "flat": tr.flat.data.clone(), "seed": tr.seed, "out": torch.stack(out)}, tmp)
    finally:
        dist.destroy_process_group()


def test_replicas_start_equal_and_in_kernel_eps_is_split_invariant(tmp_path, monkeypatch):
    """Trainer broadcasts rank 0's initial weights (the CLIs do not seed by default) and shares the eps seed; with eps
    drawn in the kernel (Philox keyed on seed, step, global image index) a 2-rank run on a ragged split reproduces the
    single-process trajectory from the same weights.  Runs the real C ABI on tests/simt_emu."""
    from tests.simt_emu.build import build
    build()
    tmp = str(tmp_path / "bc.pt")
    mp.spawn(_bcast_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    got = torch.load(tmp)
    import spatial_vae.functional as SF
    import spatial_vae.models as M
    from spatial_vae.trainer import Trainer
    from tests import emu_backend
    emu_backend.install(monkeypatch)
    torch.manual_seed(50)
    with contextlib.redirect_stdout(io.StringIO()):
        p = M.SpatialGenerator(3, 16, n_out=1, num_layers=2)
        q = M.InferenceNetwork(36, 6, 12, num_layers=2)
    tr = Trainer(p, q, SF.StepSpec(family="mnist", theta_prior=0.7, precision="parity"), lr=1e-3, seed=got["seed"])
    g = torch.Generator().manual_seed(9)
    grid = O.make_grid(6, 6)
    out = []
    for B in (5, 4):
        y = (torch.rand(B, 36, generator=g) > 0.7).float() * torch.rand(B, 36, generator=g)
        out.append(tr.step(grid, y).clone())
    np.testing.assert_allclose(got["out"].numpy(), torch.stack(out).numpy(), rtol=2e-5)
    np.testing.assert_allclose(got["flat"].numpy(), tr.flat.data.numpy(), rtol=1e-4, atol=1e-6)
