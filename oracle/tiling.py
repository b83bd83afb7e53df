"""CPU restatement (numpy index maps) of the reference's patch extraction / reassembly.

Test infrastructure only (see ``oracle/__init__.py``).  The reference implements these with
``F.pad(mode="reflect")`` + ``F.unfold`` / ``F.fold`` (src/util/tiling.py); this restatement uses
explicit index arithmetic so that it is an independent statement of the
same maps, and is pinned against the reference's outputs in ``tests/golden``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _reflect(t: np.ndarray, n: int) -> np.ndarray:
    """torch "reflect" padding index (no edge duplication): -t for t<0, 2(n-1)-t for t>=n."""
    t = np.where(t < 0, -t, t)
    return np.where(t >= n, 2 * (n - 1) - t, t)


def image_to_patches(images: np.ndarray, outer: int, inner: int) -> Tuple[np.ndarray, List[Tuple[int, int]]]:
    """src/util/tiling.py:10-64.  ``images [N,H,W]`` -> ``patches [N*nV*nH, outer, outer]`` plus the
    per-image ``(nV, nH)`` list.  ``pad=(outer-inner)//2`` reflect padding on every side (:25,:40-44),
    extra reflect padding on the bottom/right up to a multiple of ``inner`` (:33-38), unfold with
    kernel ``outer`` / stride ``inner`` (:46-50), patches row-major over ``(py, px)`` (:55-60)."""
    n, h, w = images.shape
    pad = (outer - inner) // 2
    vpad = (inner - h % inner) % inner
    hpad = (inner - w % inner) % inner
    nv, nh = (h + vpad) // inner, (w + hpad) // inner
    r = np.arange(outer)
    rows = _reflect(inner * np.arange(nv)[:, None] - pad + r[None, :], h)      # [nv, outer]
    cols = _reflect(inner * np.arange(nh)[:, None] - pad + r[None, :], w)      # [nh, outer]
    # out[n, py, px, r, s] = img[n, rows[py, r], cols[px, s]]
    out = images[:, rows[:, None, :, None], cols[None, :, None, :]]
    return np.ascontiguousarray(out.reshape(n * nv * nh, outer, outer)), [(nv, nh)] * n


def weight_matrix(tile: int) -> np.ndarray:
    """src/util/tiling.py:67-88: ``exp(-0.1 * dist_to_centre)`` evaluated in float64 (numpy scalars),
    stored element-wise into an fp32 tensor (:85), then divided by its max in fp32 (:86)."""
    c = (tile - 1) / 2
    i = np.arange(tile, dtype=np.float64)
    d = np.sqrt((i[:, None] - c) ** 2 + (i[None, :] - c) ** 2)
    w = np.exp(-0.1 * d).astype(np.float32)
    return (w / w.max()).astype(np.float32)


def _fold(tiles: np.ndarray, nv: int, nh: int, k: int, stride: int, weights) -> np.ndarray:
    """Overlap-accumulate ``tiles [nv*nh, k, k]`` (optionally times ``weights [k,k]``) into an
    ``(nv*stride, nh*stride)`` image with ``padding=(k-stride)//2`` -- what ``F.fold`` does at
    tiling.py:124-137 / :165-178 -- and divide by the accumulated weights."""
    pad = (k - stride) // 2
    big_h, big_w = nv * stride + 2 * pad, nh * stride + 2 * pad
    acc = np.zeros((big_h, big_w), dtype=np.float32)
    norm = np.zeros((big_h, big_w), dtype=np.float32)
    wts = np.ones((k, k), dtype=np.float32) if weights is None else weights
    t = tiles.reshape(nv, nh, k, k).astype(np.float32)
    # F.fold (col2im on CPU) adds the contributions of a pixel in increasing kernel-offset order (ky, kx), i.e. with
    # DEcreasing patch row, then decreasing patch column; fp32 addition is not associative, so pixels covered by four
    # patches only come out bit-identical to the reference (tests/golden/tiling.npz) in this order
    for py in range(nv - 1, -1, -1):
        for px in range(nh - 1, -1, -1):
            ys, xs = py * stride, px * stride
            acc[ys:ys + k, xs:xs + k] += t[py, px] * wts
            norm[ys:ys + k, xs:xs + k] += wts
    acc = acc[pad:pad + nv * stride, pad:pad + nh * stride]
    norm = norm[pad:pad + nv * stride, pad:pad + nh * stride]
    return (acc / norm).astype(np.float32)


def patches_to_image_weighted_average(tiles: np.ndarray, info: Sequence[Tuple[int, int]], k: int,
                                      inner: int) -> np.ndarray:
    """src/util/tiling.py:91-140 -> ``[1, nV*inner, nH*inner]`` (first image's info only, :109-112).
    Output includes the bottom/right padding area (not cropped back to H x W)."""
    nv, nh = info[0]
    return _fold(tiles[: nv * nh], nv, nh, k, inner, weight_matrix(k))[None]


def patches_to_image(tiles: np.ndarray, info: Sequence[Tuple[int, int]], k: int, inner: int) -> np.ndarray:
    """src/util/tiling.py:143-181: the same fold with unit weights."""
    nv, nh = info[0]
    return _fold(tiles[: nv * nh], nv, nh, k, inner, None)[None]


def black_mask(patches: np.ndarray) -> np.ndarray:
    """src/util/tiling.py:184-198: a patch is black iff its fp32 mean ``< 1e-10``.
    (fp32 mean of the [k,k] patch, as ``torch.Tensor.mean`` computes it.)"""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(patches))
    m = t.reshape(t.shape[0], -1).mean(dim=1)
    return (m < 1e-10).numpy()


def filter_and_remember_black_patches(patches: np.ndarray):
    """src/util/tiling.py:244-271."""
    mask = black_mask(patches)
    black = [int(i) for i in np.nonzero(mask)[0]]
    return patches[~mask], black, patches.shape


def reintegrate_black_patches(processed: np.ndarray, black: Sequence[int], original_shape) -> np.ndarray:
    """src/util/tiling.py:274-303: zeros at the black indices, processed patches elsewhere, in order."""
    full = np.zeros((original_shape[0],) + processed.shape[1:], dtype=processed.dtype)
    keep = np.ones(original_shape[0], dtype=bool)
    keep[list(black)] = False
    full[keep] = processed
    return full


def normalize_scan(scan: np.ndarray) -> np.ndarray:
    """src/util/visualization.py:113-126 (also preprocessing.py:127-137 per volume):
    ``(x - min) / (max - min)`` in fp32."""
    scan = scan.astype(np.float32)
    mn, mx = scan.min(), scan.max()
    return ((scan - mn) / (mx - mn)).astype(np.float32)
