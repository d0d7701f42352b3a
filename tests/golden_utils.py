"""Helpers shared by the oracle-pinning tests (CPU) and the CUDA parity tests (GPU)."""
import os

import torch

from oracle import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt"))


def load_case(name):
    g = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    d = getattr(synth.Dims, g["dims_factory"])(**g["dims_kwargs"])
    sd = synth.make_state_dict(d, seed=0)
    inp = synth.make_inputs(d, seed=1, **g["input_kwargs"])
    return g, d, sd, inp


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()
