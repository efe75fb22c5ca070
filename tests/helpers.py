"""Shared helpers for the tests: load golden fixtures into oracle parameter dicts."""
import os

import numpy as np
import torch

from oracle import svae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def state(d, prefix, dtype=torch.float32):
    n = len(prefix)
    return {k[n:]: torch.from_numpy(v).to(dtype) for k, v in d.items() if k.startswith(prefix)}


def oracle_params(d, dtype=torch.float32, p_prefix="p.", q_prefix="q."):
    dec = O.decoder_params_from_state(state(d, p_prefix, dtype))
    enc = O.encoder_params_from_state(state(d, q_prefix, dtype))
    return dec, enc


def golden_grads(d, dtype=torch.float32):
    """Reference gradients in optimiser order (p_net.parameters() then q_net.parameters())."""
    gp = state(d, "gp.", dtype)
    gq = state(d, "gq.", dtype)
    dec = O.decoder_params_from_state(gp)
    enc = O.encoder_params_from_state(gq)
    return O.flatten_params(dec, enc)


def option_cfg(d, family="particles"):
    """StepConfig of a particles_opt_* fixture (softplus / resid flags; expand_coords and bilinear follow from
    the parameter shapes)."""
    cfg = cfg_of(d, family)
    cfg.softplus = bool(int(d.get("opt_softplus", 0)))
    cfg.resid = bool(int(d.get("opt_resid", 0)))
    return cfg


def cfg_of(d, family):
    return O.StepConfig(
        family=family,
        rotate=bool(int(d.get("rotate", 1))), translate=bool(int(d.get("translate", 1))),
        dx_scale=float(d["dx_scale"]), theta_prior=float(d["theta_prior"]),
        z_scale=float(d.get("z_scale", 1.0)),
        activation=str(d["act"]) if "act" in d else "tanh")
