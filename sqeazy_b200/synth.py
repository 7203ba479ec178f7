"""Seeded synthetic light-sheet-like uint16 volumes, modelled on the reference's test
fixtures (tests/volume_fixtures.hpp:18-93: ellipsoid shell + noise; bench/benchmark_fixtures.hpp:64-125).

Presets
  scmos : background 100 + N(0, 3) clipped at 0, shell adds 2000        (sCMOS light-sheet like; headline)
  ref   : background 100 + Exp(mean 0.01*S), shell adds S = 0.6*65535   (the reference fixtures' own noise model)
Degenerate volumes for LZ4 extremes: zeros, random, ramp.

`numpy_volume` (CPU, tests) and `torch_volume` (device, bench at full size) share the geometry; their
noise streams differ (numpy Philox vs torch Philox), which is fine: parity is always checked on the
same bytes on both sides.
"""
from __future__ import annotations

import numpy as np


def _shell_mask_np(shape, centre_shift=0.0):
    Z, Y, X = shape
    z = np.arange(Z, dtype=np.float32)[:, None, None]
    y = np.arange(Y, dtype=np.float32)[None, :, None]
    x = np.arange(X, dtype=np.float32)[None, None, :]
    r = ((x - X / 2 - centre_shift) ** 2) / (X / 4) ** 2 + ((y - Y / 2) ** 2) / (Y / 4) ** 2 + ((z - Z / 2) ** 2) / (0.6 * Z) ** 2
    return np.abs(1.0 - r) < 0.07


def numpy_volume(shape, preset="scmos", seed=0x5EA2, index=0):
    """uint16 volume, C order {Z,Y,X}; `index` = position in a time-lapse (shell drifts 1 voxel/volume)."""
    shape = tuple(int(s) for s in shape)
    rng = np.random.Generator(np.random.Philox(seed + index))
    if preset == "zeros":
        return np.zeros(shape, dtype=np.uint16)
    if preset == "random":
        return rng.integers(0, 65536, size=shape, dtype=np.uint16)
    if preset == "ramp":
        return (np.arange(int(np.prod(shape)), dtype=np.uint64) % 32768).astype(np.uint16).reshape(shape)
    shell = _shell_mask_np(shape, float(index))
    if preset == "scmos":
        v = 100.0 + rng.normal(0.0, 3.0, size=shape).astype(np.float32)
        v = np.clip(v, 0, None) + 2000.0 * shell
    elif preset == "ref":
        S = 0.6 * 65535
        v = 100.0 + rng.exponential(0.01 * S, size=shape).astype(np.float32) + S * shell
    else:
        raise ValueError(preset)
    return np.clip(np.rint(v), 0, 65535).astype(np.uint16)


def torch_volume(shape, preset="scmos", seed=0x5EA2, index=0, device="cuda", slab=64):
    """same model generated on the device, slab by slab (full-size bench volumes never touch the host).
    Returns an int16 tensor holding the uint16 bit patterns (torch has no arithmetic on uint16)."""
    import torch

    Z, Y, X = (int(s) for s in shape)
    g = torch.Generator(device=device)
    g.manual_seed(seed + index)
    out = torch.empty((Z, Y, X), dtype=torch.int16, device=device)
    y = torch.arange(Y, dtype=torch.float32, device=device)[None, :, None]
    x = torch.arange(X, dtype=torch.float32, device=device)[None, None, :]
    ryx = ((x - X / 2 - float(index)) ** 2) / (X / 4) ** 2 + ((y - Y / 2) ** 2) / (Y / 4) ** 2
    for z0 in range(0, Z, slab):
        z1 = min(Z, z0 + slab)
        n = z1 - z0
        if preset == "zeros":
            out[z0:z1] = 0
            continue
        if preset == "random":
            v = torch.randint(0, 65536, (n, Y, X), generator=g, device=device, dtype=torch.int32)
            out[z0:z1] = v.to(torch.int16)
            continue
        z = torch.arange(z0, z1, dtype=torch.float32, device=device)[:, None, None]
        shell = ((1.0 - (ryx + ((z - Z / 2) ** 2) / (0.6 * Z) ** 2)).abs() < 0.07).to(torch.float32)
        if preset == "scmos":
            v = 100.0 + 3.0 * torch.randn((n, Y, X), generator=g, device=device)
            v = v.clamp_(min=0) + 2000.0 * shell
        elif preset == "ref":
            S = 0.6 * 65535
            e = torch.empty((n, Y, X), device=device).exponential_(1.0 / (0.01 * S), generator=g)
            v = 100.0 + e + S * shell
        else:
            raise ValueError(preset)
        out[z0:z1] = v.round_().clamp_(0, 65535).to(torch.int32).to(torch.int16)
    return out
