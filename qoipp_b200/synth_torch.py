"""The synthetic image classes of synth.py that the performance configurations use, computed with torch integer ops so
that 8K / 16K images and the 8192-image batch are generated on the device in milliseconds instead of minutes on the host.

Bit-identical to synth.py (tests/test_synth_torch.py compares them on the CPU): the same
``h(i) = splitmix64(seed ^ i * 0x9E3779B97F4A7C15)``, carried in int64 (two's complement wrap = uint64 wrap, logical
shifts and unsigned remainders done by hand).  Bench / test plumbing only: no codec work happens here.
"""
from __future__ import annotations

import torch

from . import synth

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLD = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    return (z >> k) & ((1 << (64 - k)) - 1)


def _splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _GOLD
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    return z ^ _lsr(z, 31)


def _hash_of(idx: torch.Tensor, seed) -> torch.Tensor:
    """seed: python int or an int64 tensor broadcastable against idx"""
    if not torch.is_tensor(seed):
        seed = _s64(seed)
    return _splitmix64(seed ^ (idx * _GOLD))


def _umod(z: torch.Tensor, m: int) -> torch.Tensor:
    """z taken as uint64, modulo a small m"""
    return ((_lsr(z, 1) % m) * 2 + (z & 1)) % m


def _value_noise(x, y, w, seed, cell=64):
    gx, gy = x // cell, y // cell
    fx, fy = (x % cell) * 256 // cell, (y % cell) * 256 // cell
    stride = w // cell + 2

    def lat(ix, iy):
        return _hash_of(iy * stride + ix, seed) & 0xFF

    v00, v10, v01, v11 = lat(gx, gy), lat(gx + 1, gy), lat(gx, gy + 1), lat(gx + 1, gy + 1)
    top = v00 * (256 - fx) + v10 * fx
    bot = v01 * (256 - fx) + v11 * fx
    return (top * (256 - fy) + bot * fy) >> 16


def _photo_rows(w, y0, y1, ch, seeds: torch.Tensor, device, opaque=False):
    """rows [y0, y1) of the `photo` class for every seed in `seeds` (int64 tensor [B]) -> uint8 [B, (y1-y0)*w*ch]"""
    n = (y1 - y0) * w
    i = torch.arange(y0 * w, y1 * w, dtype=torch.int64, device=device)[None, :]
    x, y = i % w, i // w
    sd = seeds.to(device)[:, None]
    hv = _hash_of(i, sd)
    out = torch.empty((seeds.numel(), n, ch), dtype=torch.uint8, device=device)
    cell_id = (y // 64) * (w // 64 + 1) + (x // 64)
    cell_h = _hash_of(cell_id, sd ^ 0x5EED)
    is_flat = _umod(cell_h, 5) == 0
    for k in range(3):
        base = _value_noise(x, y, w, sd + 101 * (k + 1))
        grain = _umod(_lsr(hv, 8 * k) if k else hv, 7) - 3
        v = torch.clamp(base + grain, 0, 255)
        v = torch.where(is_flat, _lsr(cell_h, 8 * (k + 1)) & 0xFF, v)
        out[:, :, k] = v.to(torch.uint8)
    if ch == 4:
        if opaque:
            out[:, :, 3] = 255
        else:
            field = _value_noise(x, y, w, sd ^ 0xA1FA, cell=128)
            a = torch.where(field > 176, torch.clamp(255 - (field - 176) * 3, 0, 255), torch.full_like(field, 255))
            out[:, :, 3] = a.to(torch.uint8)
    return out.reshape(seeds.numel(), n * ch)


def _noise_rows(w, y0, y1, ch, seeds, device, resync=False):
    i = torch.arange(y0 * w, y1 * w, dtype=torch.int64, device=device)[None, :]
    hv = _hash_of(i, seeds.to(device)[:, None])
    n = (y1 - y0) * w
    out = torch.empty((seeds.numel(), n, ch), dtype=torch.uint8, device=device)
    if resync:  # synth.gen_resync
        out[:, :, 0] = 0xFE if ch == 3 else 0xFF
        out[:, :, 1] = (_lsr(hv, 8) & 0xFF).to(torch.uint8)
        out[:, :, 2] = (_lsr(hv, 16) & 0xFF).to(torch.uint8)
        if ch == 4:
            out[:, :, 3] = torch.where(i % 2 == 0, 0xFF, 0xFE).to(torch.uint8)
    else:  # synth.gen_noise
        for k in range(ch):
            out[:, :, k] = ((_lsr(hv, 8 * k) if k else hv) & 0xFF).to(torch.uint8)
    return out.reshape(seeds.numel(), n * ch)


KINDS = ("photo", "photo_opaque", "noise", "resync")


def generate(kind: str, w: int, h: int, ch: int, seeds=None, device="cpu", band_pixels: int = 1 << 24) -> torch.Tensor:
    """uint8 tensor [len(seeds), w*h*ch] on `device`; seeds default to synth.py's class seed (one image).
    `photo_opaque` is `photo` with alpha 255 (RGB photo content in RGBA layout)."""
    if kind not in KINDS:
        raise ValueError(f"synth_torch has no class {kind!r}")
    base = "photo" if kind == "photo_opaque" else kind
    if seeds is None:
        seeds = [synth.BASE_SEED + synth.CLASSES.index(base)]
    st = torch.as_tensor([_s64(int(s)) for s in seeds], dtype=torch.int64)
    B = st.numel()
    out = torch.empty((B, w * h * ch), dtype=torch.uint8, device=device)
    rows = max(1, min(h, band_pixels // max(1, w * B)))  # bound the int64 temporaries
    for y0 in range(0, h, rows):
        y1 = min(h, y0 + rows)
        if base == "photo":
            band = _photo_rows(w, y0, y1, ch, st, device, opaque=kind == "photo_opaque")
        else:
            band = _noise_rows(w, y0, y1, ch, st, device, resync=base == "resync")
        out[:, y0 * w * ch: y1 * w * ch] = band
    return out
