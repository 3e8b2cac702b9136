"""Deterministic, integer-only synthetic images.

Every class is a pure function of (width, height, channels, seed) built from
``h(i) = splitmix64(seed ^ i * 0x9E3779B97F4A7C15)`` so that any host reproduces the same bytes.
Returned arrays are ``uint8`` of shape ``(height * width * channels,)`` in row-major RGB(A) order,
the layout qoipp takes (reference: source/util.hpp:319-327).
"""
from __future__ import annotations

import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
BASE_SEED = 0x51F0

CLASSES = ("noise", "flat", "flat0", "gradient", "long_runs", "photo", "dither", "palette", "hash_collide",
           "alpha_toggle", "resync", "wrap")


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + GOLD
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hashes(n: int, seed: int, start: int = 0) -> np.ndarray:
    i = np.arange(start, start + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) ^ (i * GOLD))


def _hash_of(idx: np.ndarray, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) ^ (idx.astype(np.uint64) * GOLD))


def _finish(px: np.ndarray, channels: int) -> np.ndarray:
    """px: (n, 4) uint8 -> flat (n * channels)"""
    return np.ascontiguousarray(px[:, :channels]).reshape(-1)


def _slot(px: np.ndarray) -> np.ndarray:
    p = px.astype(np.uint32)
    return (p[:, 0] * 3 + p[:, 1] * 5 + p[:, 2] * 7 + p[:, 3] * 11) & 63


def _xy(w: int, h: int):
    i = np.arange(w * h, dtype=np.int64)
    return i % w, i // w


def gen_noise(w, h, ch, seed):
    hv = hashes(w * h, seed)
    px = np.stack([(hv >> np.uint64(8 * k)).astype(np.uint8) for k in range(4)], axis=1)
    if ch == 3:
        px[:, 3] = 255
    return _finish(px, ch)


def gen_flat(w, h, ch, seed, colour=(40, 80, 120, 255)):
    px = np.tile(np.array(colour, dtype=np.uint8), (w * h, 1))
    return _finish(px, ch)


def gen_flat0(w, h, ch, seed):
    return gen_flat(w, h, ch, seed, (0, 0, 0, 255))


def _gradient_px(w, h):
    x, y = _xy(w, h)
    r = (255 * x) // max(w - 1, 1)
    g = (255 * y) // max(h - 1, 1)
    b = (255 * (x + y)) // max(w + h - 2, 1)
    a = np.full_like(r, 255)
    return np.stack([r, g, b, a], axis=1)


def gen_gradient(w, h, ch, seed):
    return _finish(_gradient_px(w, h).astype(np.uint8), ch)


def gen_dither(w, h, ch, seed):
    px = _gradient_px(w, h)
    hv = hashes(w * h, seed)
    for k in range(3):
        px[:, k] = np.clip(px[:, k] + ((hv >> np.uint64(16 * k)) % np.uint64(3)).astype(np.int64) - 1, 0, 255)
    return _finish(px.astype(np.uint8), ch)


def gen_long_runs(w, h, ch, seed):
    n = w * h
    hv = hashes(n // 2 + 16, seed)  # more runs than ever needed is fine for small n; trimmed below
    lens = (1 + (hv % np.uint64(5000))).astype(np.int64)
    need = int(np.searchsorted(np.cumsum(lens), n)) + 1
    lens = lens[:need]
    pal_h = hashes(16, seed ^ 0xABCDEF)
    pal = np.stack([(pal_h >> np.uint64(8 * k)).astype(np.uint8) for k in range(4)], axis=1)
    if ch == 3:
        pal[:, 3] = 255
    else:
        pal[:, 3] |= 0x80
    col = ((hv[:need] >> np.uint64(40)) % np.uint64(16)).astype(np.int64)
    px = np.repeat(pal[col], lens, axis=0)[:n]
    return _finish(px, ch)


def _value_noise(w, h, seed, cell=64):
    """integer value noise: hashed lattice every `cell` px, bilinear in 8.8 fixed point -> 0..255"""
    x, y = _xy(w, h)
    gx, gy = x // cell, y // cell
    fx, fy = (x % cell) * 256 // cell, (y % cell) * 256 // cell
    stride = w // cell + 2

    def lat(ix, iy):
        return (_hash_of(iy * stride + ix, seed) & np.uint64(0xFF)).astype(np.int64)

    v00, v10, v01, v11 = lat(gx, gy), lat(gx + 1, gy), lat(gx, gy + 1), lat(gx + 1, gy + 1)
    top = v00 * (256 - fx) + v10 * fx
    bot = v01 * (256 - fx) + v11 * fx
    return (top * (256 - fy) + bot * fy) >> 16


def gen_photo(w, h, ch, seed):
    n = w * h
    x, y = _xy(w, h)
    hv = hashes(n, seed)
    px = np.empty((n, 4), dtype=np.int64)
    for k in range(3):
        base = _value_noise(w, h, seed + 101 * (k + 1))
        grain = ((hv >> np.uint64(8 * k)) % np.uint64(7)).astype(np.int64) - 3
        px[:, k] = np.clip(base + grain, 0, 255)
    # flat 64x64 blocks in one of five cells
    cell_id = (y // 64) * (w // 64 + 1) + (x // 64)
    cell_h = _hash_of(cell_id, seed ^ 0x5EED)
    is_flat = (cell_h % np.uint64(5)) == 0
    for k in range(3):
        px[:, k] = np.where(is_flat, ((cell_h >> np.uint64(8 * (k + 1))) & np.uint64(0xFF)).astype(np.int64), px[:, k])
    px[:, 3] = 255
    if ch == 4:
        # soft-edged translucent blobs: alpha dips where a coarse noise field is high
        field = _value_noise(w, h, seed ^ 0xA1FA, cell=128)
        px[:, 3] = np.where(field > 176, np.clip(255 - (field - 176) * 3, 0, 255), 255)
    return _finish(px.astype(np.uint8), ch)


def _palette_distinct_slots(count, seed, ch):
    pal, used, i = [], set(), 0
    while len(pal) < count:
        hv = int(hashes(1, seed ^ 0x9A1E77E, i)[0])
        i += 1
        p = [hv & 255, (hv >> 8) & 255, (hv >> 16) & 255, 255 if ch == 3 else 0xC0 | ((hv >> 24) & 0x3F)]
        s = (p[0] * 3 + p[1] * 5 + p[2] * 7 + p[3] * 11) & 63
        if s not in used:
            used.add(s)
            pal.append(p)
    return np.array(pal, dtype=np.uint8)


def gen_palette(w, h, ch, seed):
    pal = _palette_distinct_slots(48, seed, ch)
    hv = hashes(w * h, seed)
    return _finish(pal[(hv % np.uint64(48)).astype(np.int64)], ch)


def gen_hash_collide(w, h, ch, seed):
    """A / B=A+(64,0,0,0) share a slot and never hit; C,C pairs do hit; zero and start pixels sprinkled in."""
    n = w * h
    hv = hashes(n, seed)
    a = np.stack([(hv >> np.uint64(8 * k)).astype(np.uint8) for k in range(4)], axis=1)
    grp = np.arange(n) // 8  # all eight pixels of a group derive from the group's first hash
    base = a[np.minimum(grp * 8, n - 1)].copy()
    if ch == 3:
        base[:, 3] = 255
    phase = np.arange(n) % 8
    px = base.copy()
    px[:, 0] = np.where((phase == 1) | (phase == 3), base[:, 0] + np.uint8(64), base[:, 0])  # B = A + 64 red
    other = (phase == 4) | (phase == 6)  # a different colour; phases 5 and 7 are A again: index hits
    px[:, 1] = np.where(other, base[:, 1] + np.uint8(37), px[:, 1])
    sel = (hv >> np.uint64(50)) % np.uint64(23)
    zero = np.array([0, 0, 0, 0 if ch == 4 else 255], dtype=np.uint8)
    start = np.array([0, 0, 0, 255], dtype=np.uint8)
    px[sel == 0] = zero
    px[sel == 1] = start
    lead = min(n, int(hv[0] % np.uint64(5)))  # sometimes a leading run of the start pixel
    px[:lead] = start
    if n > 3:
        px[min(n - 1, lead + 1)] = zero
    return _finish(px, ch)


def gen_alpha_toggle(w, h, ch, seed):
    px = _gradient_px(w, h)
    hv = hashes(w * h, seed)
    px[:, 3] = np.where(np.arange(w * h) % 2 == 0, 255, (hv & np.uint64(0xFF)).astype(np.int64))
    return _finish(px.astype(np.uint8), ch)


def gen_resync(w, h, ch, seed):
    """red byte pinned to the literal-op tags so a parse started one byte late never re-synchronises"""
    n = w * h
    hv = hashes(n, seed)
    px = np.empty((n, 4), dtype=np.uint8)
    px[:, 0] = 0xFE if ch == 3 else 0xFF
    px[:, 1] = (hv >> np.uint64(8)).astype(np.uint8)
    px[:, 2] = (hv >> np.uint64(16)).astype(np.uint8)
    px[:, 3] = 255 if ch == 3 else np.where(np.arange(n) % 2 == 0, 0xFF, 0xFE).astype(np.uint8)
    return _finish(px, ch)


def gen_wrap(w, h, ch, seed):
    """small steps that cross 255 -> 0 and 0 -> 255 (wrapping i8 deltas)"""
    n = w * h
    hv = hashes(n, seed)
    step = np.stack([((hv >> np.uint64(8 * k)) % np.uint64(5)).astype(np.int64) - 2 for k in range(3)], axis=1)
    big = ((hv >> np.uint64(40)) % np.uint64(11)) == 0
    step[big, 1] += ((hv[big] >> np.uint64(44)) % np.uint64(40)).astype(np.int64) - 20
    start = np.array([254, 1, 255], dtype=np.int64)
    rgb = (np.cumsum(step, axis=0) + start) & 255
    px = np.concatenate([rgb, np.full((n, 1), 255, dtype=np.int64)], axis=1)
    if ch == 4:
        px[:, 3] = np.where(((hv >> np.uint64(52)) % np.uint64(97)) == 0, 0x7F, 255)
    return _finish(px.astype(np.uint8), ch)


_GEN = {
    "noise": gen_noise, "flat": gen_flat, "flat0": gen_flat0, "gradient": gen_gradient, "long_runs": gen_long_runs,
    "photo": gen_photo, "dither": gen_dither, "palette": gen_palette, "hash_collide": gen_hash_collide,
    "alpha_toggle": gen_alpha_toggle, "resync": gen_resync, "wrap": gen_wrap,
}


def generate(kind: str, width: int, height: int, channels: int, seed: int | None = None) -> np.ndarray:
    """Synthetic image of class `kind`; default seed is 0x51F0 + class index."""
    if seed is None:
        seed = BASE_SEED + CLASSES.index(kind)
    out = _GEN[kind](width, height, channels, seed)
    assert out.dtype == np.uint8 and out.size == width * height * channels
    return out
