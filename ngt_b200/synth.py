"""Synthetic datasets of the BASELINE.json shapes.

Low-intrinsic-dimension linear-latent model: with A in R^{L x D} ~ N(0,1) (seed 12345),
    x = mu + (sigma / sqrt(L)) * z A + eta,   z ~ N(0, I_L),  eta ~ N(0, tau^2 I_D)
base rows use seed 1 (+1000*rank offsets for shards), queries seed 2. i.i.d. Gaussian data in 128-d
would make every graph index look broken; this has SIFT-like difficulty.

Host (numpy, Philox) version; `make_device` is the torch twin used for sets too large to stream
through the host. One run always feeds the SAME array to the engine and to the CPU oracle, so the two
generators never need to agree bit for bit.
"""
import numpy as np

SHAPES = {
    # name: (D, L, mu, sigma, tau, lo, hi, round)
    "sift": (128, 32, 64.0, 40.0, 4.0, 0.0, 255.0, True),     # C2 / C5
    "glove": (100, 32, 0.0, 1.0, 0.1, None, None, False),     # C3
    "gist": (960, 64, 0.25, 0.15, 0.02, 0.0, 1.0, False),     # C4
}
MIX_SEED = 12345


def _mix(shape):
    D, L = SHAPES[shape][0], SHAPES[shape][1]
    g = np.random.Generator(np.random.Philox(key=MIX_SEED))
    return g.standard_normal((L, D), dtype=np.float32)


def make(shape, n, seed, chunk=262144):
    """[n, D] float32 rows of the named shape."""
    D, L, mu, sigma, tau, lo, hi, rnd = SHAPES[shape]
    A = _mix(shape)
    g = np.random.Generator(np.random.Philox(key=seed))
    out = np.empty((n, D), np.float32)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        z = g.standard_normal((m, L), dtype=np.float32)
        eta = g.standard_normal((m, D), dtype=np.float32)
        x = mu + (sigma / np.sqrt(L)) * (z @ A) + tau * eta
        if rnd:
            x = np.rint(x)
        if lo is not None:
            x = np.clip(x, lo, hi)
        out[s:s + m] = x
    return out


def hamming_from(x, mu):
    """C5 Hamming variant: bit j = [x_j > mu], packed little-endian into D/8 bytes per row."""
    bits = (x > mu).astype(np.uint8)
    return np.packbits(bits, axis=1, bitorder="little")


def make_device(shape, n, seed, device, chunk=1 << 20, out_dtype=None):
    """torch twin of `make` (same model, torch's Philox stream) producing the rows directly in HBM."""
    import torch
    D, L, mu, sigma, tau, lo, hi, rnd = SHAPES[shape]
    A = torch.from_numpy(_mix(shape)).to(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dt = out_dtype or torch.float32
    out = torch.empty((n, D), dtype=dt, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        z = torch.randn((m, L), generator=g, device=device)
        eta = torch.randn((m, D), generator=g, device=device)
        x = mu + (sigma / L ** 0.5) * (z @ A) + tau * eta
        if rnd:
            x = torch.round(x)
        if lo is not None:
            x = torch.clamp(x, lo, hi)
        out[s:s + m] = x.to(dt)
    return out
