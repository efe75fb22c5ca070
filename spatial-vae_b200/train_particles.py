#!/usr/bin/env python
"""Train spatial-VAE on single-particle EM stacks on B200 GPUs.

Command-line compatible with the reference train_particles.py (flags/defaults at reference
train_particles.py:277-318, stdout table 'Epoch Split ELBO Error KL' at :495,514-527); hyphenated
flags as in the reference, underscore spellings accepted too.
"""
from __future__ import print_function, division

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import spatial_vae.models as models            # noqa: E402
import spatial_vae.functional as SF            # noqa: E402
import spatial_vae.ctf as C                    # noqa: E402
from spatial_vae import driver as D            # noqa: E402
from spatial_vae.trainer import Trainer        # noqa: E402

eval_minibatch = D.eval_minibatch_particles


def _unpack(mb):
    return (mb[0], mb[1]) if len(mb) > 1 else (mb[0], None)


def train_epoch(iterator, x_coord, mask, p_net, q_net, optim, rotate=True, translate=True, dx_scale=0.1,
                theta_prior=np.pi, augment_rotation=False, z_scale=1, epoch=1, num_epochs=1, N=1, use_cuda=False):
    """The reference's loop over a DataLoader of (y,) or (y, ctf) with a torch optimiser (train_particles.py:151-203)."""
    def call(mb):
        y, ctf = _unpack(mb)
        return eval_minibatch(x_coord, y, mask, ctf, p_net, q_net, rotate=rotate, translate=translate,
                              dx_scale=dx_scale, theta_prior=theta_prior, augment_rotation=augment_rotation,
                              z_scale=z_scale, use_cuda=use_cuda) + (None,)
    return D.epoch_loop(iterator, call, train=True, p_net=p_net, q_net=q_net, optim=optim, epoch=epoch,
                        num_epochs=num_epochs, total=N)


def eval_model(iterator, x_coord, mask, p_net, q_net, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
               z_scale=1, use_cuda=False):
    """train_particles.py:206-248 (no augmentation at evaluation time)."""
    def call(mb):
        y, ctf = _unpack(mb)
        return eval_minibatch(x_coord, y, mask, ctf, p_net, q_net, rotate=rotate, translate=translate,
                              dx_scale=dx_scale, theta_prior=theta_prior, z_scale=z_scale, use_cuda=use_cuda) + (None,)
    return D.epoch_loop(iterator, call, train=False, p_net=p_net, q_net=q_net)


def _both(name):
    return ['--' + name, '--' + name.replace('-', '_')] if '-' in name else ['--' + name]


def parse(argv=None):
    p = argparse.ArgumentParser('Train spatial-VAE on particle datasets')
    p.add_argument('train_path', nargs='?', help='path to training data')
    p.add_argument('test_path', nargs='?', help='path to testing data')
    p.add_argument(*_both('ctf-train'), dest='ctf_train', help='path to CTF parameters for training images')
    p.add_argument(*_both('ctf-test'), dest='ctf_test', help='path to CTF parameters for testing images')
    p.add_argument('--scale', default=1, type=float)
    p.add_argument('-z', *_both('z-dim'), dest='z_dim', type=int, default=2)
    p.add_argument(*_both('p-hidden-dim'), dest='p_hidden_dim', type=int, default=500)
    p.add_argument(*_both('p-num-layers'), dest='p_num_layers', type=int, default=2)
    p.add_argument(*_both('q-hidden-dim'), dest='q_hidden_dim', type=int, default=500)
    p.add_argument(*_both('q-num-layers'), dest='q_num_layers', type=int, default=2)
    p.add_argument('-a', '--activation', choices=['tanh', 'relu'], default='tanh')
    p.add_argument('--softplus', action='store_true')
    p.add_argument('--resid', action='store_true')
    p.add_argument(*_both('expand-coords'), dest='expand_coords', action='store_true')
    p.add_argument('--bilinear', action='store_true')
    p.add_argument(*_both('fit-noise'), dest='fit_noise', action='store_true')
    p.add_argument('--vanilla', action='store_true')
    p.add_argument(*_both('no-rotate'), dest='no_rotate', action='store_true')
    p.add_argument(*_both('no-translate'), dest='no_translate', action='store_true')
    p.add_argument(*_both('dx-scale'), dest='dx_scale', type=float, default=0.1)
    p.add_argument(*_both('theta-prior'), dest='theta_prior', type=float, default=np.pi)
    p.add_argument('-l', *_both('learning-rate'), dest='learning_rate', type=float, default=1e-4)
    p.add_argument(*_both('minibatch-size'), dest='minibatch_size', type=int, default=100)
    p.add_argument(*_both('augment-rotation'), dest='augment_rotation', action='store_true')
    p.add_argument(*_both('z-delay'), dest='z_delay', type=int, default=0)
    p.add_argument('--normalize', action='store_true')
    p.add_argument('-c', '--crop', type=int, default=-1)
    p.add_argument(*_both('save-prefix'), dest='save_prefix')
    p.add_argument(*_both('save-interval'), dest='save_interval', default=10, type=int)
    p.add_argument(*_both('num-epochs'), dest='num_epochs', type=int, default=100)
    p.add_argument('-d', '--device', type=int, default=-2)
    p.add_argument(*_both('no-preload'), dest='no_preload', action='store_true')
    p.add_argument('--mask', action='store_true')
    D.add_b200_flags(p, hyphen=True)
    return p.parse_args(argv)


load_images = D.load_particle_stack


def main(argv=None):
    args = parse(argv)
    if args.fit_noise and args.ctf_train is not None:
        raise SystemExit('--fit-noise cannot be combined with CTF filtering (the reference crashes on it, '
                         'train_particles.py:121-124,137)')
    device = D.pick_device(args.device)
    rank = D.init_distributed(device)
    if args.seed is not None:
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
    digits = int(np.log10(args.num_epochs)) + 1

    if args.synthetic > 0:
        side = args.synthetic_size or 40
        g = np.random.default_rng(1234)
        images_train = g.standard_normal((args.synthetic, side, side)).astype(np.float32)
        images_test = g.standard_normal((max(args.synthetic // 4, 1), side, side)).astype(np.float32)
    else:
        images_train, images_test = load_images(args.train_path), load_images(args.test_path)
    print('# train:', images_train.shape, ', test:', images_test.shape, file=sys.stderr)
    if args.crop > 0:
        from spatial_vae.image import crop
        images_train, images_test = crop(images_train, args.crop), crop(images_test, args.crop)
        print('# cropped to:', args.crop, file=sys.stderr)
    n, m = images_train.shape[1:]
    if args.normalize:
        print('# normalizing particles', file=sys.stderr)
        images_train, images_test = D.normalize_particles(images_train), D.normalize_particles(images_test)

    # CTF kernels are odd-sized: 40x40 images get 39x39 kernels (reference train_particles.py:353-358); they are built
    # on the device, all particles in one launch (svae_ctf_filter), not in the reference's per-particle numpy loop
    kn, km = (n - 1 if n % 2 == 0 else n), (m - 1 if m % 2 == 0 else m)
    ctf_train = ctf_test = None
    if args.ctf_train is not None:
        print('# loading CTF filters:', args.ctf_train, file=sys.stderr)
        ctf_train = SF.ctf_filter(C.parse_ctf(args.ctf_train), kn, km, scale=args.scale, device=device)
    if args.ctf_test is not None:
        print('# loading CTF filters:', args.ctf_test, file=sys.stderr)
        ctf_test = SF.ctf_filter(C.parse_ctf(args.ctf_test), kn, km, scale=args.scale, device=device)

    x_coord = D.make_grid(n, m, device)
    y_train = torch.from_numpy(np.ascontiguousarray(images_train)).float().view(-1, n * m).to(device)
    y_test = torch.from_numpy(np.ascontiguousarray(images_test)).float().view(-1, n * m).to(device)
    mask = None
    if args.mask:   # circular mask (reference train_particles.py:387-396)
        print('# masking particles', file=sys.stderr)
        mask = D.circular_mask(n, m).to(device)
        print('# masking to size:', int(mask.sum()), file=sys.stderr)

    print('# training with z-dim:', args.z_dim, file=sys.stderr)
    activation = D.activation_from_flag(args.activation, 'particles')
    if args.vanilla:      # reference train_particles.py:425-431: standard MLP generator, no rotation / translation inference
        print('# using the vanilla MLP generator architecture', file=sys.stderr)
        p_net = models.VanillaGenerator(n * m, args.z_dim, args.p_hidden_dim, n_out=2 if args.fit_noise else 1,
                                        num_layers=args.p_num_layers, activation=activation,
                                        softplus=args.softplus).to(device)
        q_net = models.InferenceNetwork(n * m, args.z_dim, args.q_hidden_dim, num_layers=args.q_num_layers,
                                        activation=activation).to(device)
        D.train_vanilla('particles', args, x_coord, y_train, y_test, p_net, q_net,
                        header='\t'.join(['Epoch', 'Split', 'ELBO', 'Error', 'KL']), rank=rank, ctf_train=ctf_train,
                        ctf_test=ctf_test, mask=mask)
        return
    print('# using the spatial generator architecture', file=sys.stderr)
    rotate, translate = not args.no_rotate, not args.no_translate
    inf_dim = args.z_dim + (1 if rotate else 0) + (2 if translate else 0)
    p_net = models.SpatialGenerator(args.z_dim, args.p_hidden_dim, n_out=2 if args.fit_noise else 1,
                                    num_layers=args.p_num_layers, activation=activation,
                                    softplus=args.softplus, resid=args.resid, expand_coords=args.expand_coords,
                                    bilinear=args.bilinear).to(device)
    q_net = models.InferenceNetwork(n * m, inf_dim, args.q_hidden_dim, num_layers=args.q_num_layers,
                                    activation=activation, resid=args.resid).to(device)
    print('# using priors: theta={}, dx={}'.format(args.theta_prior, args.dx_scale), file=sys.stderr)

    spec = SF.StepSpec(family='particles', rotate=rotate, translate=translate, dx_scale=args.dx_scale,
                       theta_prior=args.theta_prior, activation=p_net.activation_code, softplus=args.softplus,
                       precision=args.precision, resid=args.resid)
    trainer = Trainer(p_net, q_net, spec, lr=args.learning_rate)
    shuffle_gen = torch.Generator().manual_seed(args.seed if args.seed is not None else 0)
    augment = None
    if args.augment_rotation and rotate:
        augment = lambda y: D._augment(y, True, True)      # host PIL rotation, as in the reference

    if rank == 0:
        print('\t'.join(['Epoch', 'Split', 'ELBO', 'Error', 'KL']))
    for epoch in range(args.num_epochs):
        z_scale = 0.0 if epoch < args.z_delay else 1.0
        e, err, kl = D.run_epoch(trainer, x_coord, y_train, train=True, minibatch_size=args.minibatch_size,
                                 generator=shuffle_gen, ctf=ctf_train, mask=mask, augment=augment, z_scale=z_scale,
                                 epoch=epoch, num_epochs=args.num_epochs)
        if rank == 0:
            print('\t'.join([str(epoch + 1), 'train', str(e), str(err), str(kl)]), flush=True)
        e, err, kl = D.run_epoch(trainer, x_coord, y_test, train=False, minibatch_size=args.minibatch_size,
                                 ctf=ctf_test, mask=mask, z_scale=z_scale)
        if rank == 0:
            print('\t'.join([str(epoch + 1), 'test', str(e), str(err), str(kl)]), flush=True)
            if args.save_prefix is not None and (epoch + 1) % args.save_interval == 0:
                D.save_models(args.save_prefix, str(epoch + 1).zfill(digits), None, p_net, q_net, device)


if __name__ == '__main__':
    main()
