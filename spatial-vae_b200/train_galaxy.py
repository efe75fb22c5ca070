#!/usr/bin/env python
"""Train spatial-VAE on RGB galaxy images on B200 GPUs.

Command-line compatible with the reference train_galaxy.py (flags/defaults at reference
train_galaxy.py:300-341, stdout table 'Epoch / ELBO / BCE loss / KL'); underscore flags as in the
reference, hyphenated spellings accepted too.
"""
from __future__ import print_function, division

import argparse
import datetime
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import spatial_vae.models as models            # noqa: E402
import spatial_vae.functional as SF            # noqa: E402
from spatial_vae import driver as D            # noqa: E402
from spatial_vae.trainer import Trainer        # noqa: E402

eval_minibatch = D.eval_minibatch_galaxy
random_minibatch_generator = D.random_minibatch_generator


def minibatch_for_display(x, y, q_net, p_net, rotate=True, translate=True, z_scale=1, use_cuda=False):
    """Same signature as the reference's galaxy driver, which takes q_net BEFORE p_net (train_galaxy.py:131), unlike
    the mnist driver (train_mnist.py:93)."""
    return D.minibatch_for_display(x, y, p_net, q_net, rotate=rotate, translate=translate, z_scale=z_scale,
                                   use_cuda=use_cuda)


def train_epoch(iterator, x_coord, p_net, q_net, optim, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
                augment_rotation=False, z_scale=1, epoch=1, num_epochs=1, train_images_len=1, use_cuda=False):
    """The reference's loop over a DataLoader with a torch optimiser (train_galaxy.py:186-231)."""
    call = lambda mb: eval_minibatch(x_coord, mb[0], p_net, q_net, rotate=rotate, translate=translate,
                                     dx_scale=dx_scale, theta_prior=theta_prior, augment_rotation=augment_rotation,
                                     z_scale=z_scale, use_cuda=use_cuda)
    return D.epoch_loop(iterator, call, train=True, p_net=p_net, q_net=q_net, optim=optim, epoch=epoch,
                        num_epochs=num_epochs, total=train_images_len)


def eval_model(iterator, x_coord, p_net, q_net, z_dim, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
               z_scale=1, use_cuda=False, to_save_image_samples=False, image_dims=None, epoch='0',
               output_dir='outputs', save_label=''):
    """train_galaxy.py:234-295: validation means; optionally PNG grids (reconstruction on the unrotated grid, the
    decoder output, and samples from the prior) of the first minibatch."""
    call = lambda mb: eval_minibatch(x_coord, mb[0], p_net, q_net, rotate=rotate, translate=translate,
                                     dx_scale=dx_scale, theta_prior=theta_prior, z_scale=z_scale, use_cuda=use_cuda)

    def dump(mb, y_hat):
        y = mb[0]
        y_display = minibatch_for_display(x_coord, y, q_net, p_net, rotate=rotate, translate=translate,
                                          z_scale=z_scale, use_cuda=use_cuda)
        y_random = random_minibatch_generator(x_coord, y, p_net, z_dim, z_scale=z_scale, use_cuda=use_cuda)
        for data, tag in ((y_display, '_dis'), (y_hat.detach(), ''), (y_random, '_rnd')):
            D.export_batch_as_image(data, '{}/images/{}{}_{}.png'.format(output_dir, epoch, tag, save_label), image_dims)

    return D.epoch_loop(iterator, call, train=False, p_net=p_net, q_net=q_net,
                        first_batch_hook=dump if (to_save_image_samples and image_dims) else None)


def _both(name):
    return ['--' + name, '--' + name.replace('_', '-')] if '_' in name else ['--' + name]


def galaxy_arguments(argv=None):
    p = argparse.ArgumentParser('Train spatial-VAE on galaxy datasets')
    p.add_argument('train_path', nargs='?', help='path to training data')
    p.add_argument('test_path', nargs='?', help='path to testing data')
    p.add_argument('-z', *_both('z_dim'), dest='z_dim', type=int, default=2)
    p.add_argument(*_both('p_hidden_dim'), dest='p_hidden_dim', type=int, default=500)
    p.add_argument(*_both('p_num_layers'), dest='p_num_layers', type=int, default=2)
    p.add_argument(*_both('q_hidden_dim'), dest='q_hidden_dim', type=int, default=5000)
    p.add_argument(*_both('q_num_layers'), dest='q_num_layers', type=int, default=2)
    p.add_argument('-a', '--activation', choices=['tanh', 'relu', 'leakyrelu', 'sigmoid'], default='tanh')
    p.add_argument('--vanilla', action='store_true')
    p.add_argument(*_both('no_rotate'), dest='no_rotate', action='store_true')
    p.add_argument(*_both('no_translate'), dest='no_translate', action='store_true')
    p.add_argument(*_both('dx_scale'), dest='dx_scale', type=float, default=0.1)
    p.add_argument(*_both('theta_prior'), dest='theta_prior', type=float, default=np.pi)
    p.add_argument('-l', *_both('learning_rate'), dest='learning_rate', type=float, default=1e-4)
    p.add_argument(*_both('minibatch_size'), dest='minibatch_size', type=int, default=100)
    p.add_argument(*_both('augment_rotation'), dest='augment_rotation', action='store_true')
    p.add_argument(*_both('z_delay'), dest='z_delay', type=int, default=0)
    p.add_argument(*_both('save_prefix'), dest='save_prefix')
    p.add_argument(*_both('save_interval'), dest='save_interval', default=10, type=int)
    p.add_argument(*_both('num_epochs'), dest='num_epochs', type=int, default=100)
    p.add_argument('-d', '--device', type=int, default=-2)
    p.add_argument(*_both('num_train_images'), dest='num_train_images', type=int, default=0)
    p.add_argument(*_both('val_split'), dest='val_split', type=int, default=50)
    p.add_argument(*_both('make_mono'), dest='make_mono', action='store_true')
    p.add_argument(*_both('logging_level'), dest='logging_level', type=str, default='INFO')
    p.add_argument(*_both('invert_colours'), dest='invert_colours', action='store_true')
    D.add_b200_flags(p)
    return p.parse_args(argv)


def main(argv=None):
    args = galaxy_arguments(argv)
    device = D.pick_device(args.device)
    rank = D.init_distributed(device)
    if args.seed is not None:
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
    out_dir = trained_dir = None
    digits = int(np.log10(args.num_epochs)) + 1
    start_time = datetime.datetime.now()
    if args.save_prefix is not None and rank == 0:
        start_time, out_dir, trained_dir, _, digits = D.prepare_output_dir(args, assume_yes=args.yes)

    if args.synthetic > 0:
        side = args.synthetic_size or 64
        g = np.random.default_rng(1234)
        images_train = (g.random((args.synthetic, side, side, 3)) * 255).astype(np.uint8)
        images_val = (g.random((max(args.synthetic // 4, 1), side, side, 3)) * 255).astype(np.uint8)
    else:
        # validation runs on the TEST file: this fork commented the --val_split split of the training file out
        # ("revert back to using test set for validation", reference train_galaxy.py:362-382)
        images_train = np.load(args.train_path)
        images_val = np.load(args.test_path)
        if args.make_mono:      # only the training set is converted (reference train_galaxy.py:366-370)
            images_train = np.mean(images_train, axis=3)
        # every data-parallel rank must hold the SAME dataset order (run_epoch hands each rank a slice of one shared
        # permutation): shuffle with a seeded generator whenever there are several ranks or --seed is given; a single
        # unseeded process shuffles with the global numpy RNG as the reference does (train_galaxy.py:372)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.seed is not None:
            np.random.default_rng(args.seed if args.seed is not None else 0).shuffle(images_train)
        else:
            np.random.shuffle(images_train)
    if args.num_train_images > 0:
        images_train = images_train[:args.num_train_images]
        images_val = images_val[:args.num_train_images]
    if images_train.ndim == 3:
        images_train = images_train[..., None]
    rows, cols, channels = images_train.shape[1:4]
    y_train = torch.from_numpy(np.ascontiguousarray(images_train)).float() / 255
    y_val = torch.from_numpy(np.ascontiguousarray(images_val)).float() / 255
    if args.invert_colours:
        y_train, y_val = 1 - y_train, 1 - y_val
    y_train = y_train.reshape(-1, rows * cols, channels).to(device)
    y_val = y_val.reshape(-1, rows * cols, channels).to(device)       # (the reference views the 3-channel test file
    #                                                                    with the training set's channel count)
    x_coord = D.make_grid(rows, cols, device)

    print('# training with z-dim:', args.z_dim, file=sys.stderr)
    activation = D.activation_from_flag(args.activation, 'galaxy')
    if args.vanilla:      # reference train_galaxy.py: standard MLP generator, no rotation / translation inference
        print('# using the vanilla MLP generator architecture', file=sys.stderr)
        p_net = models.VanillaGenerator(rows * cols, args.z_dim, args.p_hidden_dim, n_out=channels,
                                        num_layers=args.p_num_layers, activation=activation).to(device)
        q_net = models.InferenceNetwork(channels * rows * cols, args.z_dim, args.q_hidden_dim,
                                        num_layers=args.q_num_layers, activation=activation).to(device)
        D.train_vanilla('galaxy', args, x_coord, y_train, y_val, p_net, q_net,
                        header='\t'.join(['Epoch', 'ELBO', 'BCE loss', 'KL']), rank=rank)
        return
    print('# using the spatial generator architecture', file=sys.stderr)
    rotate, translate = not args.no_rotate, not args.no_translate
    inf_dim = args.z_dim + (1 if rotate else 0) + (2 if translate else 0)
    p_net = models.SpatialGenerator(args.z_dim, args.p_hidden_dim, n_out=channels, num_layers=args.p_num_layers,
                                    activation=activation).to(device)
    q_net = models.InferenceNetwork(channels * rows * cols, inf_dim, args.q_hidden_dim,
                                    num_layers=args.q_num_layers, activation=activation).to(device)
    print('# using priors: theta={}, dx={}'.format(args.theta_prior, args.dx_scale), file=sys.stderr)

    spec = SF.StepSpec(family='galaxy', rotate=rotate, translate=translate, dx_scale=args.dx_scale,
                       theta_prior=args.theta_prior, activation=p_net.activation_code, precision=args.precision)
    trainer = Trainer(p_net, q_net, spec, lr=args.learning_rate)
    shuffle_gen = torch.Generator().manual_seed(args.seed if args.seed is not None else 0)
    augment = None
    if args.augment_rotation and rotate:
        augment = lambda y: D._augment(y, True, True, channels)

    header = '\t'.join(['Epoch', 'ELBO', 'BCE loss', 'KL'])
    if rank == 0:
        print(header)
    train_lines, val_lines = [header], [header]
    for epoch in range(args.num_epochs):
        z_scale = 0.0 if epoch < args.z_delay else 1.0
        e, err, kl = D.run_epoch(trainer, x_coord, y_train, train=True, minibatch_size=args.minibatch_size,
                                 generator=shuffle_gen, augment=augment, z_scale=z_scale, epoch=epoch,
                                 num_epochs=args.num_epochs)
        line = '\t'.join(map(str, [epoch, e, err, kl]))
        train_lines.append(line)
        if rank == 0:
            print(line, flush=True)
        if len(y_val) > 0:
            hook = D.sample_dump_hook(args, out_dir, epoch, (rows, cols), x_coord, p_net, q_net, rotate, translate,
                                      z_scale) if rank == 0 else None
            e, err, kl = D.run_epoch(trainer, x_coord, y_val, train=False, minibatch_size=args.minibatch_size,
                                     z_scale=z_scale, first_batch_hook=hook)
            line = '\t'.join(map(str, [epoch, e, err, kl]))
            val_lines.append(line)
            if rank == 0:
                print(line, flush=True)
    if rank == 0:
        D.save_models(args.save_prefix, str(args.num_epochs).zfill(digits), trained_dir, p_net, q_net, device)
        if out_dir:
            D.write_results(out_dir, train_lines, val_lines)
        end = datetime.datetime.now()
        print(f"End : {end.strftime('%y%m%d_%H%M%S')}")
        print(f"Elapsed time: {end - start_time}")


if __name__ == '__main__':
    main()
