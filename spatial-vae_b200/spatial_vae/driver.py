"""Shared machinery of the three command lines (train_mnist.py, train_particles.py, train_galaxy.py).

The reference triplicates eval_minibatch / train_epoch / eval_model in each script
(train_mnist.py:24-226, train_particles.py:22-245, train_galaxy.py:27-294); here the three scripts
are thin flag definitions over this module.  The per-minibatch work is one fused call
(spatial_vae.functional.elbo_step or, in the epoch loops, spatial_vae.trainer.Trainer.step); the
batch-size-weighted running means the reference keeps on the host with three .item() syncs per
step (train_mnist.py:152-164) are kept on the device and read once per progress update.
"""
from __future__ import annotations

import datetime
import math
import os
import sys
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import functional as SF
from . import models
from .trainer import Trainer, shard_bounds


# --------------------------------------------------------------------------------------------------
# eval_minibatch: same signatures and return values as the reference functions
# --------------------------------------------------------------------------------------------------
def _spec_for(family, p_net, rotate, translate, dx_scale, theta_prior, z_scale=1.0, precision=None):
    return SF.StepSpec(family=family, rotate=bool(rotate), translate=bool(translate), dx_scale=float(dx_scale),
                       theta_prior=float(theta_prior), z_scale=float(z_scale),
                       activation=getattr(p_net, "activation_code", 0), softplus=bool(getattr(p_net, "softplus", False)),
                       resid=bool(getattr(p_net, "resid", False)),
                       precision=precision or getattr(p_net, "precision", None) or SF.default_precision())


def _is_spatial(p_net):
    return isinstance(p_net, models.SpatialGenerator)


def eval_minibatch_mnist(x, y, p_net, q_net, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
                         use_cuda=False, eps=None):
    """train_mnist.eval_minibatch (reference train_mnist.py:24-90) -> (elbo, log_p_x_g_z, kl_div, y_hat)."""
    if use_cuda:
        y = y.cuda()
    spec = _spec_for("mnist", p_net, rotate, translate, dx_scale, theta_prior)
    elbo, logp, kl, y_hat, _ = SF.elbo_step(spec, x, y, p_net, q_net, eps=eps, want_y_hat=True)
    return elbo, logp, kl, y_hat.view(y.size(0), -1)


def rotate_images_bicubic(y, n, offsets, channels=None):
    """The reference's augmentation (train_particles.py:36-43 / train_galaxy.py:44-54) rotates every image of the
    minibatch with PIL on the host (device -> host -> device round trip plus a Python loop).  Here the same
    arithmetic (Pillow's bicubic affine resampling, uint8 round trip for RGB) runs in one kernel on the device."""
    deg = 360 * np.asarray(offsets, dtype=np.float64) / 2 / np.pi
    return SF.rotate_bicubic(y, n, n, deg, channels=channels or 1, quantize_u8=channels is not None)


def _augment(y, rotate, augment_rotation, channels=None):
    b = y.size(0)
    n = int(np.sqrt(y.size(1)))
    if not (rotate and augment_rotation):
        return None, None
    offset = np.random.uniform(0, 2 * np.pi, size=b)       # unseeded numpy RNG, as in the reference
    if rotate < 1:
        offset *= np.random.binomial(1, p=rotate, size=b)
    y_rot = rotate_images_bicubic(y, n, offset, channels)
    return y_rot, torch.from_numpy(offset).float().to(y.device)


def eval_minibatch_particles(x, y, mask, ctf, p_net, q_net, rotate=True, translate=True, dx_scale=0.1,
                             theta_prior=np.pi, augment_rotation=False, z_scale=1, use_cuda=False, eps=None):
    """train_particles.eval_minibatch (reference train_particles.py:22-148) -> (elbo, log_p_x_g_z, kl_div)."""
    y_rot, offset = _augment(y, rotate, augment_rotation)
    if use_cuda:
        y = y.cuda()
        y_rot = y_rot.cuda() if y_rot is not None else None
    if ctf is not None and p_net.layers[-2].out_features > 1:
        # the reference's variance convolution lacks groups= and crashes here (train_particles.py:121-124,137)
        raise RuntimeError("CTF filtering cannot be combined with --fit-noise (the reference raises as well)")
    spec = _spec_for("particles", p_net, rotate, translate, dx_scale, theta_prior, z_scale)
    elbo, logp, kl, _, _ = SF.elbo_step(spec, x, y, p_net, q_net, eps=eps, y_enc=y_rot, theta_offset=offset,
                                        ctf=ctf.reshape(ctf.size(0), ctf.size(-2), ctf.size(-1)) if ctf is not None else None,
                                        mask=mask)
    return elbo, logp, kl


def eval_minibatch_galaxy(x, y, p_net, q_net, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
                          augment_rotation=False, z_scale=1, use_cuda=False, eps=None):
    """train_galaxy.eval_minibatch (reference train_galaxy.py:27-128) -> (elbo, log_p_x_g_z, kl_div, y_hat)."""
    channels = y.size(2)
    y_rot, offset = _augment(y, rotate, augment_rotation, channels)
    if use_cuda:
        y = y.cuda()
        y_rot = y_rot.cuda() if y_rot is not None else None
    spec = _spec_for("galaxy", p_net, rotate, translate, dx_scale, theta_prior, z_scale)
    elbo, logp, kl, y_hat, _ = SF.elbo_step(spec, x, y, p_net, q_net, eps=eps, y_enc=y_rot, theta_offset=offset,
                                            want_y_hat=True)
    return elbo, logp, kl, y_hat.view(y.size(0), -1, channels)


def minibatch_for_display(x, y, p_net, q_net, rotate=True, translate=True, z_scale=1, use_cuda=False):
    """Decode the sampled z on the UNrotated grid (reference train_mnist.py:93-124, train_galaxy.py:131-163)."""
    batch = y.size(0)
    if use_cuda:
        y = y.cuda()
    with torch.no_grad():
        z_mu, z_logstd = q_net(y.reshape(batch, -1))
        r = torch.empty_like(z_mu).normal_()
        z = torch.exp(z_logstd) * r + z_mu
        if rotate:
            z = z[:, 1:]
        if translate:
            z = z[:, 2:]
        z = z * z_scale
        y_hat = p_net(x.expand(batch, x.size(0), x.size(1)).contiguous(), z.contiguous())
    return y_hat.view(batch, -1) if y.dim() == 2 else y_hat.view(batch, -1, y.size(2))


def random_minibatch_generator(x, y, p_net, z_dim, z_scale=1, use_cuda=False):
    """Decode z ~ N(0, I) (reference train_galaxy.py:166-183)."""
    batch = y.size(0)
    with torch.no_grad():
        z = torch.empty(batch, z_dim, device=x.device).normal_() * z_scale
        y_hat = p_net(x.expand(batch, x.size(0), x.size(1)).contiguous(), z)
    return y_hat.view(batch, -1, y.size(2)) if y.dim() == 3 else y_hat.view(batch, -1)


# --------------------------------------------------------------------------------------------------
# --vanilla: the non-spatial MLP generator (reference models.py:135-172, train_mnist.py:351-357).  NOT on the fused
# path: the generator and the loss are plain PyTorch modules / autograd (the encoder still runs the library's
# kernels behind autograd), trained by the reference's own loop with torch.optim.Adam.
# --------------------------------------------------------------------------------------------------
def vanilla_eval_minibatch(family, x, y, p_net, q_net, z_scale=1.0, ctf=None, mask=None):
    """The reference eval_minibatch with rotate = translate = False (train_mnist.py:24-90, train_particles.py:22-148,
    train_galaxy.py:27-128), evaluated with PyTorch ops -> (elbo, log_p_x_g_z, kl_div, y_hat)."""
    import torch.nn.functional as F
    b = y.size(0)
    z_mu, z_logstd = q_net(y.reshape(b, -1))
    z_std = torch.exp(z_logstd)
    z = z_std * torch.empty_like(z_mu).normal_() + z_mu
    kl_div = torch.sum(-z_logstd + 0.5 * z_std ** 2 + 0.5 * z_mu ** 2 - 0.5, 1).mean()
    y_hat = p_net(x.expand(b, x.size(0), x.size(1)), z * z_scale)
    if family in ("mnist", "galaxy"):
        yv = y.reshape(b, -1)
        log_p = -F.binary_cross_entropy(y_hat.reshape(b, -1), yv) * yv.size(1)
    else:
        params = y_hat.reshape(b, -1)
        yv = y.reshape(b, -1)
        P = yv.size(1)
        mu, logvar = (params[:, :P], params[:, P:]) if params.size(1) > P else (params, None)
        if ctf is not None:
            n = int(np.sqrt(P))
            mu = F.conv2d(mu.reshape(1, b, n, n), ctf.reshape(b, 1, ctf.size(-2), ctf.size(-1)), padding=ctf.size(-1) // 2,
                          groups=b).reshape(b, -1)
        if mask is not None:
            sel = mask.reshape(-1).bool()
            mu, yv = mu[:, sel], yv[:, sel]
            logvar = logvar[:, sel] if logvar is not None else None
        if logvar is not None:
            log_p = (-0.5 * torch.sum((mu - yv) ** 2 / torch.exp(logvar) + logvar, 1)).mean()
        else:
            log_p = (-0.5 * torch.sum((mu - yv) ** 2, 1)).mean()
    return log_p - kl_div, log_p, kl_div, y_hat


def train_vanilla(family, args, x_coord, y_train, y_test, p_net, q_net, *, header, rank=0, ctf_train=None, ctf_test=None,
                  mask=None):
    """--vanilla training: the reference's train_epoch / eval_model loops over DataLoaders with torch.optim.Adam."""
    print('# --vanilla: non-spatial MLP generator; plain PyTorch modules and autograd, OUTSIDE the fused B200 path',
          file=sys.stderr)
    optim = torch.optim.Adam(list(p_net.parameters()) + list(q_net.parameters()), lr=args.learning_rate)
    def loader(y, c, shuffle):
        tensors = (y,) if c is None else (y, c)
        return torch.utils.data.DataLoader(torch.utils.data.TensorDataset(*tensors), batch_size=args.minibatch_size,
                                           shuffle=shuffle)
    z_scale = 1.0
    def call(mb):
        e, lp, kl, yh = vanilla_eval_minibatch(family, x_coord, mb[0], p_net, q_net, z_scale=z_scale,
                                               ctf=mb[1] if len(mb) > 1 else None, mask=mask)
        return e, lp, kl, yh
    if rank == 0:
        print(header)
    for epoch in range(args.num_epochs):
        e, err, kl = epoch_loop(loader(y_train, ctf_train, True), call, train=True, p_net=p_net, q_net=q_net, optim=optim,
                                epoch=epoch, num_epochs=args.num_epochs, total=y_train.size(0))
        if rank == 0:
            print('\t'.join(map(str, [epoch, e, err, kl])), flush=True)
        with torch.no_grad():
            e, err, kl = epoch_loop(loader(y_test, ctf_test, False), call, train=False, p_net=p_net, q_net=q_net,
                                    total=y_test.size(0))
        if rank == 0:
            print('\t'.join(map(str, [epoch, e, err, kl])), flush=True)
    return p_net, q_net


# --------------------------------------------------------------------------------------------------
# epoch loops
# --------------------------------------------------------------------------------------------------
def make_grid(n_rows, n_cols, device=None):
    """(P,2) pixel coordinates, x in [-1,1] along columns, y from +1 down to -1 along rows
    (reference train_mnist.py:316-320)."""
    xs = np.linspace(-1, 1, n_cols)
    ys = np.linspace(1, -1, n_rows)
    x0, x1 = np.meshgrid(xs, ys)
    g = torch.from_numpy(np.stack([x0.ravel(), x1.ravel()], 1)).float()
    return g.to(device) if device is not None else g


class RunningMeans:
    """Batch-size-weighted running means of (elbo, error, kl) kept on the device
    (the reference's host arithmetic at train_mnist.py:156-164)."""

    def __init__(self, device):
        self.sums = torch.zeros(3, dtype=torch.float64, device=device)
        self.count = 0

    def update(self, means: torch.Tensor, n: int):
        self.sums += means.double() * n
        self.count += n

    def read(self):
        m = (self.sums / max(self.count, 1)).tolist()
        return m[0], -m[1], m[2]      # elbo, error = -log p(x|z), kl


def epoch_permutation(n, shuffle, generator):
    return torch.randperm(n, generator=generator) if shuffle else torch.arange(n)


def run_epoch(trainer: Trainer, x_coord, data, *, train: bool, minibatch_size: int, generator=None, ctf=None,
              mask=None, augment=None, z_scale=None, epoch=0, num_epochs=1, progress=True, first_batch_hook=None):
    """One pass over `data` (N, ...) resident on the device.  Every rank walks the same permutation
    (same CPU generator seed) and takes its contiguous slice of each minibatch; the last minibatch may
    be ragged (DataLoader(drop_last=False), reference train_mnist.py:395)."""
    N = data.shape[0]
    dev = data.device
    perm = epoch_permutation(N, train, generator)
    means = RunningMeans(dev)
    world, rank = trainer.world, trainer.rank
    for start in range(0, N, minibatch_size):
        idx = perm[start:start + minibatch_size]
        bsz = idx.numel()
        lo, hi = shard_bounds(bsz, world, rank)
        idx_local = idx[lo:hi].to(dev)
        y = SF.gather_rows(data, idx_local)
        c = SF.gather_rows(ctf, idx_local) if ctf is not None else None
        if train:
            y_enc = theta_offset = None
            if augment is not None:
                y_enc, theta_offset = augment(y)
            # one CUDA-graph replay per minibatch (captured once per batch shape, Trainer.step_graphed; the NCCL
            # allreduce of a multi-rank step is part of the captured graph); SVAE_NO_GRAPH=1 enqueues kernel by kernel
            step = trainer.step if os.environ.get("SVAE_NO_GRAPH") == "1" else trainer.step_graphed
            res = step(x_coord, y, global_batch=bsz, y_enc=y_enc, theta_offset=theta_offset, ctf=c, mask=mask,
                       z_scale=z_scale, image_offset=lo)
        else:
            want = first_batch_hook is not None and start == 0
            res, y_hat = trainer.evaluate(x_coord, y, global_batch=bsz, ctf=c, mask=mask, want_y_hat=want,
                                          z_scale=z_scale)
            if want:
                first_batch_hook(y, y_hat)
        means.update(res, bsz)
        if progress and train and rank == 0 and (start // minibatch_size) % 20 == 0:
            e, err, kl = means.read()
            line = '# [{}/{}] training {:.1%}, ELBO={:.5f}, Error={:.5f}, KL={:.5f}'.format(
                epoch + 1, num_epochs, means.count / N, e, err, kl)
            print(line, end='\r', file=sys.stderr)
    if progress and train and rank == 0:
        print(' ' * 80, end='\r', file=sys.stderr)
    return means.read()


# --------------------------------------------------------------------------------------------------
# the reference's own per-epoch loops (train_epoch / eval_model of the three drivers), for callers that bring their
# DataLoader and torch optimiser: each minibatch goes through eval_minibatch (the fused step behind autograd),
# loss.backward(), optim.step(), optim.zero_grad(), and the running means are accumulated as the reference does
# (train_mnist.py:127-171,174-226; train_particles.py:151-203,206-248; train_galaxy.py:186-231,234-295).
# The command lines of this build use run_epoch + Trainer instead (flat buffers, fused Adam, CUDA graph).
# --------------------------------------------------------------------------------------------------
def epoch_loop(iterator, eval_call, *, train, p_net, q_net, optim=None, epoch=1, num_epochs=1, total=1,
               first_batch_hook=None):
    """eval_call(minibatch) -> (elbo, log_p_x_g_z, kl_div, extras); returns (elbo, loss, kl) running means as Python
    floats, weighted by minibatch size."""
    p_net.train(train)
    q_net.train(train)
    count = 0
    elbo_accum = loss_accum = kl_accum = 0.0
    for it, mb in enumerate(iterator):
        elbo, logp, kl, extras = eval_call(mb)
        b = mb[0].size(0)
        if train:
            (-elbo).backward()
            optim.step()
            optim.zero_grad()
        elbo_v, loss_v, kl_v = elbo.item(), -logp.item(), kl.item()
        count += b
        loss_accum += b * (loss_v - loss_accum) / count
        elbo_accum += b * (elbo_v - elbo_accum) / count
        kl_accum += b * (kl_v - kl_accum) / count
        if train:
            line = '# [{}/{}] training {:.1%}, ELBO={:.5f}, Error={:.5f}, KL={:.5f}'.format(
                epoch + 1, num_epochs, count / total, elbo_accum, loss_accum, kl_accum)
            print(line, end='\r', file=sys.stderr)
        elif it == 0 and first_batch_hook is not None:
            first_batch_hook(mb, extras)
    if train:
        print(' ' * 80, end='\r', file=sys.stderr)
    return elbo_accum, loss_accum, kl_accum


# --------------------------------------------------------------------------------------------------
# run bookkeeping (reference src/misc_tools.py, src/file_tools.py): thin, behaviour compatible
# --------------------------------------------------------------------------------------------------
def activation_from_flag(name, script):
    """'relu' means LeakyReLU in mnist/particles (train_mnist.py:344-348) but true ReLU in galaxy, whose
    'leakyrelu' choice silently stays Tanh (typo at train_galaxy.py:429)."""
    if script in ("mnist", "particles"):
        return nn.Tanh if name == "tanh" else nn.LeakyReLU
    return {"tanh": nn.Tanh, "relu": nn.ReLU, "sigmoid": nn.Sigmoid}.get(name, nn.Tanh)


def pick_device(d):
    """-d flag (reference train_mnist.py:323-327): -1 = CPU (not supported here), -2 = current CUDA device."""
    if d == -1 or not torch.cuda.is_available():
        raise SystemExit("this build of spatial-VAE runs on NVIDIA B200 GPUs only (no CPU path); "
                         "use the reference implementation for CPU runs")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if d >= 0:
        local = d
        print('# using CUDA device:', d, file=sys.stderr)
    torch.cuda.set_device(local)
    return torch.device("cuda", local)


def init_distributed(device):
    """One process per GPU under torchrun; single process otherwise."""
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
        import atexit
        atexit.register(lambda: dist.is_initialized() and dist.destroy_process_group())
    return dist.get_rank() if dist.is_initialized() else 0


def prepare_output_dir(args, assume_yes=False):
    """outputs_<prefix>/{trained,images} wiped and recreated, command.txt written
    (reference misc_tools.py:48-74).  The reference blocks on input(); --yes or a non-tty skips it."""
    out_existing = 'outputs_{}'.format(args.save_prefix)
    if not assume_yes:
        if sys.stdin.isatty():
            if input('WARNING Will clear the outputs directory if it exists. Continue (y/n and Enter)?').lower() == 'n':
                raise SystemExit(0)
        elif os.path.isdir(out_existing):
            # the reference always prompts; a batch job must say --yes before an existing directory is deleted
            raise SystemExit(f"{out_existing} exists and stdin is not a terminal: pass --yes to clear it")
    import shutil
    start = datetime.datetime.now()
    print(f"Start : {start.strftime('%y%m%d_%H%M%S')}")
    out = 'outputs_{}'.format(args.save_prefix)
    trained, images = os.path.join(out, 'trained'), os.path.join(out, 'images')
    if os.path.isdir(out):
        shutil.rmtree(out)
    for d in (out, trained, images):
        os.makedirs(d, exist_ok=True)
    with open(os.path.join(out, 'command.txt'), 'w') as f:
        for k, v in vars(args).items():
            print(f'{k}: {v}', file=f)
    digits = int(np.log10(args.num_epochs)) + 1
    return start, out, trained, images, digits


def save_label(args):
    """reference misc_tools.py:15-28"""
    names = {'z_dim': 'z', 'p_num_layers': 'pnl', 'q_num_layers': 'qnl', 'num_layers': 'nl', 'num_epochs': 'ep'}
    label = str(args.save_prefix) + '_'
    for k, v in vars(args).items():
        if k in names:
            label += names[k] + str(v)
    return label


def save_models(path_prefix, epoch_str, trained_dir, p_net, q_net, device):
    """Whole-module pickles *_generator_epochN.sav / *_inference_epochN.sav, saved from CPU in eval mode
    (reference misc_tools.py:87-104, train_particles.py:530-543).  The saved copies are detached
    clones so the trainer's flat parameter buffer stays on the device."""
    import copy
    if path_prefix is None:
        return
    for net, tag in ((p_net, 'generator'), (q_net, 'inference')):
        path = path_prefix + '_{}_epoch{}.sav'.format(tag, epoch_str)
        if trained_dir:
            path = os.path.join(trained_dir, path)
        clone = copy.deepcopy(net).eval().cpu()
        torch.save(clone, path)


def export_batch_as_image(data, output, image_dims, channels_last=True):
    """PNG grid of a batch (reference misc_tools.py:30-39); needs torchvision."""
    try:
        from torchvision.utils import save_image
    except Exception:      # torchvision is optional in this build
        return
    images = data.view(data.size(0), *image_dims, -1)
    if channels_last:
        images = images.permute(0, 3, 1, 2)
    save_image(images.cpu(), output, nrow=int(data.size(0) ** 0.5), padding=3, pad_value=0.5)


def sample_dump_hook(args, out_dir, epoch, image_dims, x_coord, p_net, q_net, rotate, translate, z_scale=1):
    """first_batch_hook for run_epoch(train=False): every save_interval epochs the reference's eval_model writes the
    first validation batch's reconstruction ({epoch}_{label}.png) and its decode on the un-rotated grid
    ({epoch}_dis_{label}.png) under outputs_<prefix>/images (train_mnist.py:213-224, train_galaxy.py:280-293)."""
    if out_dir is None or args.save_interval <= 0 or (epoch + 1) % args.save_interval != 0:
        return None
    label = save_label(args)

    def hook(y, y_hat):
        shown = minibatch_for_display(x_coord, y, p_net, q_net, rotate=rotate, translate=translate, z_scale=z_scale)
        export_batch_as_image(shown, '{}/images/{}_dis_{}.png'.format(out_dir, epoch, label), image_dims)
        export_batch_as_image(y_hat.detach().reshape(shown.shape), '{}/images/{}_{}.png'.format(out_dir, epoch, label),
                              image_dims)
    return hook


def add_b200_flags(parser, hyphen=False):
    """Flags this build adds on top of the reference's."""
    sep = '-' if hyphen else '_'
    parser.add_argument('--precision', choices=['fast', 'parity_tc', 'parity'], default='fast',
                        help='fast: bf16 tcgen05 hidden GEMMs with fp32 accumulation; parity_tc: 3-term bf16 splits on '
                             'tcgen05 (fp32 accuracy); parity: fp32 FFMA everywhere')
    parser.add_argument('--seed', type=int, default=None, help='seed torch (and the shuffling) for reproducible runs')
    parser.add_argument('--yes', action='store_true', help='do not prompt before clearing the outputs directory')
    parser.add_argument(f'--synthetic', type=int, default=0,
                        help='train on this many synthetic images of the dataset\'s shape instead of loading files')
    parser.add_argument(f'--synthetic{sep}size', type=int, default=0, help='image side for --synthetic')


def normalize_particles(images):
    """Per-particle zero mean / unit std over the pixels (reference train_particles.py:339-347)."""
    n, m = images.shape[1:]
    flat = images.reshape(-1, n * m)
    mu, std = flat.mean(1), flat.std(1)
    return (images - mu[:, np.newaxis, np.newaxis]) / std[:, np.newaxis, np.newaxis]


def circular_mask(n, m):
    """Boolean (n*m,) mask of the pixels closer than min(n,m)/2 to (n/2, m/2) (reference train_particles.py:387-396)."""
    yy, xx = np.ogrid[:n, :m]
    dist = np.sqrt((n / 2 - yy) ** 2 + (m / 2 - xx) ** 2)
    return (torch.from_numpy(dist) < min(n, m) / 2).view(-1)


def load_particle_stack(path):
    """.mrc/.mrcs through spatial_vae.mrc, .npy through numpy (reference train_particles.py:248-255)."""
    if path.endswith('mrc') or path.endswith('mrcs'):
        from . import mrc
        with open(path, 'rb') as f:
            images, _, _ = mrc.parse(f.read())
        return images
    return np.load(path)


def write_results(output_dir, train_lines, val_lines):
    with open(os.path.join(output_dir, 'train.txt'), 'w') as f:
        print('\n'.join(train_lines), file=f)
    with open(os.path.join(output_dir, 'val.txt'), 'w') as f:
        print('\n'.join(val_lines), file=f)
