#!/usr/bin/env python
"""Train spatial-VAE on (rotated / translated) MNIST-like images on B200 GPUs.

Command-line compatible with the reference train_mnist.py (flags and defaults at reference
train_mnist.py:232-263, stdout table 'Epoch / ELBO / BCE loss / KL' at :406,423-446); both the
underscore spellings the reference parses and the hyphenated ones its README shows are accepted.
Run under torchrun for data-parallel training (one process per GPU, one NCCL allreduce per step).
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

eval_minibatch = D.eval_minibatch_mnist
minibatch_for_display = D.minibatch_for_display


def train_epoch(iterator, x_coord, p_net, q_net, optim, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
                epoch=1, num_epochs=1, N=1, use_cuda=False):
    """The reference's loop over a DataLoader with a torch optimiser (train_mnist.py:127-171)."""
    call = lambda mb: eval_minibatch(x_coord, mb[0], p_net, q_net, rotate=rotate, translate=translate,
                                     dx_scale=dx_scale, theta_prior=theta_prior, use_cuda=use_cuda)
    return D.epoch_loop(iterator, call, train=True, p_net=p_net, q_net=q_net, optim=optim, epoch=epoch,
                        num_epochs=num_epochs, total=N)


def eval_model(iterator, x_coord, p_net, q_net, rotate=True, translate=True, dx_scale=0.1, theta_prior=np.pi,
               use_cuda=False, to_save_image_samples=False, image_dims=None, epoch='0', output_dir='outputs',
               save_label=''):
    """train_mnist.py:174-226: validation means; optionally PNG grids of the first minibatch."""
    call = lambda mb: eval_minibatch(x_coord, mb[0], p_net, q_net, rotate=rotate, translate=translate,
                                     dx_scale=dx_scale, theta_prior=theta_prior, use_cuda=use_cuda)

    def dump(mb, y_hat):
        y_display = minibatch_for_display(x_coord, mb[0], p_net, q_net, rotate=rotate, translate=translate,
                                          use_cuda=use_cuda)
        D.export_batch_as_image(y_display, '{}/images/{}_dis_{}.png'.format(output_dir, epoch, save_label), image_dims)
        D.export_batch_as_image(y_hat.detach(), '{}/images/{}_{}.png'.format(output_dir, epoch, save_label), image_dims)

    return D.epoch_loop(iterator, call, train=False, p_net=p_net, q_net=q_net,
                        first_batch_hook=dump if (to_save_image_samples and image_dims) else None)


def _both(name):
    return ['--' + name, '--' + name.replace('_', '-')] if '_' in name else ['--' + name]


def mnist_arguments(argv=None):
    p = argparse.ArgumentParser('Train spatial-VAE on MNIST datasets')
    p.add_argument('--dataset', choices=['mnist', 'mnist-rotated', 'mnist-rotated-translated', 'galaxy'],
                   default='mnist-rotated-translated',
                   help='which MNIST datset to train/validate on (default: mnist-rotated-translated)')
    p.add_argument('-z', *_both('z_dim'), dest='z_dim', type=int, default=2, help='latent variable dimension (default: 2)')
    p.add_argument(*_both('p_hidden_dim'), dest='p_hidden_dim', type=int, default=500)
    p.add_argument(*_both('q_hidden_dim'), dest='q_hidden_dim', type=int, default=500)
    p.add_argument(*_both('num_layers'), dest='num_layers', type=int, default=2, help='number of hidden layers (default: 2)')
    p.add_argument('-a', '--activation', choices=['tanh', 'relu'], default='tanh')
    p.add_argument('--vanilla', action='store_true', help='standard MLP generator (not on the B200 fused path)')
    p.add_argument(*_both('no_rotate'), dest='no_rotate', action='store_true')
    p.add_argument(*_both('no_translate'), dest='no_translate', action='store_true')
    p.add_argument(*_both('dx_scale'), dest='dx_scale', type=float, default=0.1)
    p.add_argument(*_both('theta_prior'), dest='theta_prior', type=float, default=np.pi / 4)
    p.add_argument('-l', *_both('learning_rate'), dest='learning_rate', type=float, default=1e-4)
    p.add_argument(*_both('minibatch_size'), dest='minibatch_size', type=int, default=100)
    p.add_argument(*_both('save_prefix'), dest='save_prefix')
    p.add_argument(*_both('save_interval'), dest='save_interval', default=10, type=int)
    p.add_argument(*_both('num_epochs'), dest='num_epochs', type=int, default=100)
    p.add_argument('-d', '--device', type=int, default=-2, help='compute device to use')
    p.add_argument(*_both('num_train_images'), dest='num_train_images', type=int, default=0)
    p.add_argument(*_both('val_split'), dest='val_split', type=int, default=50)
    D.add_b200_flags(p)
    return p.parse_args(argv)


def load_images(args):
    if args.synthetic > 0:
        n = args.synthetic_size or 28
        g = np.random.default_rng(1234)
        tr = ((g.random((args.synthetic, n, n)) > 0.8) * g.random((args.synthetic, n, n)) * 255).astype(np.uint8)
        te = ((g.random((max(args.synthetic // 4, 1), n, n)) > 0.8) * 255 * g.random((max(args.synthetic // 4, 1), n, n))).astype(np.uint8)
        return tr, te
    if args.dataset == 'mnist':
        import torchvision
        print('# training on MNIST', file=sys.stderr)
        tr = torchvision.datasets.MNIST('data/mnist/', train=True, download=True).data.numpy()
        te = torchvision.datasets.MNIST('data/mnist/', train=False, download=True).data.numpy()
        return tr, te
    if args.dataset == 'mnist-rotated':
        print('# training on rotated MNIST', file=sys.stderr)
        base = 'data/mnist_rotated'
    elif args.dataset == 'galaxy':
        print('# training on mono-chromed galaxy_zoo', file=sys.stderr)
        return (np.mean(np.load('data/galaxy_zoo/galaxy_zoo_train.npy'), axis=3),
                np.mean(np.load('data/galaxy_zoo/galaxy_zoo_test.npy'), axis=3))
    else:
        print('# training on rotated and translated MNIST', file=sys.stderr)
        base = 'data/mnist_rotated_translated'
    return np.load(base + '/images_train.npy'), np.load(base + '/images_test.npy')


def main(argv=None):
    args = mnist_arguments(argv)
    device = D.pick_device(args.device)
    rank = D.init_distributed(device)
    if args.seed is not None:
        torch.manual_seed(args.seed)
    out_dir = trained_dir = None
    digits = int(np.log10(args.num_epochs)) + 1
    start_time = datetime.datetime.now()
    if args.save_prefix is not None and rank == 0:
        start_time, out_dir, trained_dir, _, digits = D.prepare_output_dir(args, assume_yes=args.yes)

    images_train, images_test = load_images(args)
    if args.num_train_images > 0:
        images_train = images_train[:args.num_train_images]
    n, m = images_train.shape[1:3]
    y_train = (torch.from_numpy(np.ascontiguousarray(images_train)).float() / 255).view(-1, n * m).to(device)
    y_test = (torch.from_numpy(np.ascontiguousarray(images_test)).float() / 255).view(-1, n * m).to(device)
    x_coord = D.make_grid(n, m, device)

    z_dim = args.z_dim
    print('# training with z-dim:', z_dim, file=sys.stderr)
    activation = D.activation_from_flag(args.activation, 'mnist')
    if args.vanilla:      # reference train_mnist.py:351-357: standard MLP generator, no rotation / translation inference
        print('# using the vanilla MLP generator architecture', file=sys.stderr)
        p_net = models.VanillaGenerator(n * m, z_dim, args.p_hidden_dim, num_layers=args.num_layers,
                                        activation=activation).to(device)
        q_net = models.InferenceNetwork(n * m, z_dim, args.q_hidden_dim, num_layers=args.num_layers,
                                        activation=activation).to(device)
        D.train_vanilla('mnist', args, x_coord, y_train, y_test, p_net, q_net,
                        header='\t'.join(['Epoch', 'ELBO', 'BCE loss', 'KL']), rank=rank)
        if rank == 0:
            D.save_models(args.save_prefix, str(args.num_epochs).zfill(digits), trained_dir, p_net, q_net, device)
        return
    print('# using the spatial generator architecture', file=sys.stderr)
    rotate, translate = not args.no_rotate, not args.no_translate
    inf_dim = z_dim + (1 if rotate else 0) + (2 if translate else 0)
    if rotate:
        print('# spatial-VAE with rotation inference', file=sys.stderr)
    if translate:
        print('# spatial-VAE with translation inference', file=sys.stderr)
    p_net = models.SpatialGenerator(z_dim, args.p_hidden_dim, n_out=1, num_layers=args.num_layers,
                                    activation=activation).to(device)
    q_net = models.InferenceNetwork(n * m, inf_dim, args.q_hidden_dim, num_layers=args.num_layers,
                                    activation=activation).to(device)
    if out_dir:
        with open(os.path.join(out_dir, 'models.txt'), 'w') as f:
            print(p_net, file=f)
            print(q_net, file=f)
    print('# using priors: theta={}, dx={}'.format(args.theta_prior, args.dx_scale), file=sys.stderr)

    spec = SF.StepSpec(family='mnist', rotate=rotate, translate=translate, dx_scale=args.dx_scale,
                       theta_prior=args.theta_prior, activation=p_net.activation_code, precision=args.precision)
    trainer = Trainer(p_net, q_net, spec, lr=args.learning_rate)
    shuffle_gen = torch.Generator().manual_seed(args.seed if args.seed is not None else 0)

    header = '\t'.join(['Epoch', 'ELBO', 'BCE loss', 'KL'])
    if rank == 0:
        print(header)
    train_lines, val_lines = [header], [header]
    for epoch in range(args.num_epochs):
        e, err, kl = D.run_epoch(trainer, x_coord, y_train, train=True, minibatch_size=args.minibatch_size,
                                 generator=shuffle_gen, epoch=epoch, num_epochs=args.num_epochs)
        line = '\t'.join(map(str, [epoch, e, err, kl]))
        train_lines.append(line)
        if rank == 0:
            print(line, flush=True)
        hook = D.sample_dump_hook(args, out_dir, epoch, (n, m), x_coord, p_net, q_net, rotate, translate) if rank == 0 else None
        e, err, kl = D.run_epoch(trainer, x_coord, y_test, train=False, minibatch_size=args.minibatch_size,
                                 first_batch_hook=hook)
        line = '\t'.join(map(str, [epoch, e, err, kl]))
        val_lines.append(line)
        if rank == 0:
            print(line, flush=True)
    if rank == 0:
        # the reference saves once, after the last epoch (train_mnist.py:448-451)
        D.save_models(args.save_prefix, str(args.num_epochs).zfill(digits), trained_dir, p_net, q_net, device)
        if out_dir:
            D.write_results(out_dir, train_lines, val_lines)
        end = datetime.datetime.now()
        print(f"End : {end.strftime('%y%m%d_%H%M%S')}")
        print(f"Elapsed time: {end - start_time}")


if __name__ == '__main__':
    main()
