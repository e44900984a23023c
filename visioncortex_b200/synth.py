"""Synthetic inputs for the BASELINE.json configurations.

Integer-only, stateless-hash based, index-addressable: the same (W, H, seed) always gives the same image.
  G1  binary noise                                  -> C1 (BinaryImage::to_clusters)
  G2  jittered-grid Voronoi "RGBA scene" with noise -> C2 / C4 (Runner)
  G3  posterised Voronoi, optional key-colour cells -> C3 / C5
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        x = x ^ (x >> np.uint64(31))
    return x


def _h(seed: int, stream: int, i: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = _mix(np.array([np.uint64(seed) ^ (np.uint64(stream) << np.uint64(56))], dtype=np.uint64))
        return _mix(base + i.astype(np.uint64))


def g1_binary(width: int, height: int, p: float, seed: int) -> np.ndarray:
    """bool (h, w): bit i set iff (h(seed,1,i) >> 32) < floor(p * 2^32)."""
    n = width * height
    thr = np.uint64(int(p * (1 << 32)))
    bits = (_h(seed, 1, np.arange(n, dtype=np.uint64)) >> np.uint64(32)) < thr
    return bits.reshape(height, width)


def _voronoi_ids(width: int, height: int, seed: int, cell: int):
    """Nearest jittered-grid site among the 3x3 neighbouring cells (squared distance, ties -> smaller id).
    Evaluated in row bands with 32-bit arithmetic (distances < 2^31 for cell <= 10000) to stay cache resident."""
    gx = (width + cell - 1) // cell
    gy = (height + cell - 1) // cell
    ids = np.arange(gx * gy, dtype=np.uint64)
    cx = (ids % np.uint64(gx)).astype(np.int64)
    cy = (ids // np.uint64(gx)).astype(np.int64)
    sx = (cx * cell + (_h(seed, 2, ids) % np.uint64(cell)).astype(np.int64)).astype(np.int32)
    sy = (cy * cell + (_h(seed, 3, ids) % np.uint64(cell)).astype(np.int64)).astype(np.int32)
    out = np.empty((height, width), dtype=np.int32)
    xs = np.arange(width, dtype=np.int32)[None, :]
    pcx = xs // cell
    band = max(1, min(height, (1 << 21) // max(1, width)))
    big = np.int32(np.iinfo(np.int32).max)
    for y0 in range(0, height, band):
        y1 = min(height, y0 + band)
        ys = np.arange(y0, y1, dtype=np.int32)[:, None]
        pcy = ys // cell
        best_d = np.full((y1 - y0, width), big, dtype=np.int32)
        best_id = np.zeros((y1 - y0, width), dtype=np.int32)
        for dy in (-1, 0, 1):
            qy = pcy + dy
            oky = (qy >= 0) & (qy < gy)
            for dx in (-1, 0, 1):
                qx = pcx + dx
                ok = oky & ((qx >= 0) & (qx < gx))
                qid = np.where(ok, qy * gx + qx, 0).astype(np.int32)
                ddx = xs - sx[qid]
                ddy = ys - sy[qid]
                dist = np.where(ok, ddx * ddx + ddy * ddy, big)
                better = (dist < best_d) | ((dist == best_d) & (qid < best_id))
                best_d = np.where(better, dist, best_d)
                best_id = np.where(better, qid, best_id)
        out[y0:y1] = best_id
    return out, gx * gy


def g2_scene(width: int, height: int, seed: int, cell: int = 48, amp: int = 3) -> np.ndarray:
    """uint8 (h, w, 4) RGBA: Voronoi cells, diagonal shading, +-amp per-channel noise, alpha 255."""
    best_id, ncell = _voronoi_ids(width, height, seed, cell)
    col24 = (_h(seed, 4, np.arange(ncell, dtype=np.uint64)) & np.uint64(0xFFFFFF)).astype(np.int64)
    ys, xs = np.mgrid[0:height, 0:width].astype(np.int64)
    shade = ((xs + 2 * ys) >> 5) & 15
    out = np.empty((height, width, 4), dtype=np.uint8)
    pix = (ys * width + xs).astype(np.uint64)
    for ch in range(3):
        base = (col24[best_id] >> (8 * ch)) & 0xFF
        v = base + shade
        if amp > 0:
            noise = (_h(seed, 5, pix * np.uint64(4) + np.uint64(ch)) % np.uint64(2 * amp + 1)).astype(np.int64) - amp
            v = v + noise
        out[..., ch] = np.clip(v, 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def g3_posterised(width: int, height: int, seed: int, cell: int = 96, key=None) -> np.ndarray:
    """uint8 (h, w, 4): G2 geometry, no noise/shading, channels quantised to {0,85,170,255};
    with ``key`` (r,g,b,a) every cell with h(seed,6,id)%10==0 is filled with the key colour."""
    best_id, ncell = _voronoi_ids(width, height, seed, cell)
    ids = np.arange(ncell, dtype=np.uint64)
    col24 = (_h(seed, 4, ids) & np.uint64(0xFFFFFF)).astype(np.int64)
    out = np.empty((height, width, 4), dtype=np.uint8)
    for ch in range(3):
        base = (col24 >> (8 * ch)) & 0xFF
        q = (base >> 6) * 85
        out[..., ch] = q[best_id].astype(np.uint8)
    out[..., 3] = 255
    if key is not None:
        keyed = (_h(seed, 6, ids) % np.uint64(10)) == 0
        m = keyed[best_id]
        out[m] = np.array(key, dtype=np.uint8)
    return out


def random_blocky(rng: np.random.Generator, width: int, height: int, ncolors: int = 3, step: int = 24,
                  block: int = 2) -> np.ndarray:
    """Small adversarial test images: low-cardinality blocky noise so that every stage-1 quirk fires."""
    bh, bw = (height + block - 1) // block, (width + block - 1) // block
    pal = rng.integers(0, 256 // step, size=(ncolors, 3)) * step
    idx = rng.integers(0, ncolors, size=(bh, bw))
    img = pal[idx].repeat(block, axis=0).repeat(block, axis=1)[:height, :width]
    out = np.empty((height, width, 4), dtype=np.uint8)
    out[..., :3] = img.astype(np.uint8)
    out[..., 3] = 255
    return out


# ---------------------------------------------------------------------------------------------------------------
# The same generators on a torch device (bench.py builds its inputs on the GPU: 512 distinct 1080p scenes or one
# 16384 x 16384 image take seconds there and minutes in numpy).  torch has no uint64 arithmetic: the hash runs in
# wrapping int64 with logical shifts and an unsigned modulo spelled out.  tests/test_synth.py checks bit equality
# with the numpy generators above.
def _t_lsr(x, k: int):
    return (x >> k) & ((1 << (64 - k)) - 1)


def _t_i64(v: int) -> int:
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _t_mix(x):
    x = x + _t_i64(0x9E3779B97F4A7C15)
    x = (x ^ _t_lsr(x, 30)) * _t_i64(0xBF58476D1CE4E5B9)
    x = (x ^ _t_lsr(x, 27)) * _t_i64(0x94D049BB133111EB)
    return x ^ _t_lsr(x, 31)


def _t_h(seed: int, stream: int, i):
    import torch

    base = _t_mix(torch.tensor([_t_i64(seed ^ (stream << 56))], dtype=torch.int64, device=i.device))
    return _t_mix(base + i.to(torch.int64))


def _t_umod(x, m: int):
    """x mod m for x read as an unsigned 64-bit value (m < 2^31)."""
    lo = x & 0x7FFFFFFFFFFFFFFF
    top = _t_lsr(x, 63)
    return (lo % m + top * ((1 << 63) % m)) % m


def _t_voronoi_ids(width: int, height: int, seed: int, cell: int, device):
    import torch

    gx = (width + cell - 1) // cell
    gy = (height + cell - 1) // cell
    ids = torch.arange(gx * gy, dtype=torch.int64, device=device)
    cx, cy = ids % gx, ids // gx
    sx = (cx * cell + _t_umod(_t_h(seed, 2, ids), cell)).to(torch.int32)
    sy = (cy * cell + _t_umod(_t_h(seed, 3, ids), cell)).to(torch.int32)
    out = torch.empty((height, width), dtype=torch.int32, device=device)
    xs = torch.arange(width, dtype=torch.int32, device=device)[None, :]
    pcx = xs // cell
    band = max(1, min(height, (1 << 24) // max(1, width)))
    big = 2 ** 31 - 1
    for y0 in range(0, height, band):
        y1 = min(height, y0 + band)
        ys = torch.arange(y0, y1, dtype=torch.int32, device=device)[:, None]
        pcy = ys // cell
        best_d = torch.full((y1 - y0, width), big, dtype=torch.int32, device=device)
        best_id = torch.zeros((y1 - y0, width), dtype=torch.int32, device=device)
        for dy in (-1, 0, 1):
            qy = pcy + dy
            oky = (qy >= 0) & (qy < gy)
            for dx in (-1, 0, 1):
                qx = pcx + dx
                ok = oky & ((qx >= 0) & (qx < gx))
                qid = torch.where(ok, qy * gx + qx, torch.zeros_like(qx + qy)).to(torch.int32)
                q64 = qid.to(torch.int64)
                ddx = xs - sx[q64]
                ddy = ys - sy[q64]
                dist = torch.where(ok, ddx * ddx + ddy * ddy, torch.full_like(ddx, big))
                better = (dist < best_d) | ((dist == best_d) & (qid < best_id))
                best_d = torch.where(better, dist, best_d)
                best_id = torch.where(better, qid, best_id)
        out[y0:y1] = best_id
    return out, gx * gy


def g2_scene_torch(width: int, height: int, seed: int, device, cell: int = 48, amp: int = 3):
    """g2_scene on a torch device: uint8 (h, w, 4) tensor."""
    import torch

    best_id, ncell = _t_voronoi_ids(width, height, seed, cell, device)
    col24 = _t_h(seed, 4, torch.arange(ncell, dtype=torch.int64, device=device)) & 0xFFFFFF
    ys = torch.arange(height, dtype=torch.int64, device=device)[:, None]
    xs = torch.arange(width, dtype=torch.int64, device=device)[None, :]
    shade = ((xs + 2 * ys) >> 5) & 15
    pix = ys * width + xs
    out = torch.empty((height, width, 4), dtype=torch.uint8, device=device)
    cid = best_id.to(torch.int64)
    for ch in range(3):
        v = ((col24[cid] >> (8 * ch)) & 0xFF) + shade
        if amp > 0:
            v = v + _t_umod(_t_h(seed, 5, pix * 4 + ch), 2 * amp + 1) - amp
        out[..., ch] = v.clamp(0, 255).to(torch.uint8)
    out[..., 3] = 255
    return out


def g3_posterised_torch(width: int, height: int, seed: int, device, cell: int = 96, key=None):
    """g3_posterised on a torch device: uint8 (h, w, 4) tensor."""
    import torch

    best_id, ncell = _t_voronoi_ids(width, height, seed, cell, device)
    ids = torch.arange(ncell, dtype=torch.int64, device=device)
    col24 = _t_h(seed, 4, ids) & 0xFFFFFF
    out = torch.empty((height, width, 4), dtype=torch.uint8, device=device)
    cid = best_id.to(torch.int64)
    for ch in range(3):
        q = (((col24 >> (8 * ch)) & 0xFF) >> 6) * 85
        out[..., ch] = q[cid].to(torch.uint8)
    out[..., 3] = 255
    if key is not None:
        keyed = _t_umod(_t_h(seed, 6, ids), 10) == 0
        out[keyed[cid]] = torch.tensor(list(key), dtype=torch.uint8, device=device)
    return out


def g1_binary_torch(width: int, height: int, p: float, seed: int, device):
    import torch

    n = width * height
    bits = _t_lsr(_t_h(seed, 1, torch.arange(n, dtype=torch.int64, device=device)), 32) < int(p * (1 << 32))
    return bits.reshape(height, width)
