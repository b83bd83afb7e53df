"""Drop-in for the inference functions of the reference's ``src/util/tiling.py`` (same names, argument
meaning and return values), running on the coalesced HBM kernels of ``libmrinr.so``.

The three training-side helpers (``filter_black_patches``, ``filter_black_patches_indices``, ``extract_center_batch``;
tiling.py:201-241,306-322) are mirrored too, so that ``src/train/training.py:18-25`` resolves under the import overlay.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import ops

_weight_cache: Dict[Tuple[int, str], torch.Tensor] = {}


def generate_weight_matrix(tile_size: int) -> torch.Tensor:
    """tiling.py:67-88: ``exp(-0.1 * distance to the tile centre)`` evaluated in float64, stored as fp32, then
    divided by its maximum in fp32.  Host computation (K*K values, once per tile size)."""
    centre = (tile_size - 1) / 2
    i = np.arange(tile_size, dtype=np.float64) - centre
    dist = np.sqrt(i[:, None] ** 2 + i[None, :] ** 2)
    w = torch.from_numpy(np.exp(-0.1 * dist).astype(np.float32))
    return w / w.max()


def _weights_on(tile_size: int, device) -> torch.Tensor:
    key = (tile_size, str(device))
    w = _weight_cache.get(key)
    if w is None:
        w = generate_weight_matrix(tile_size).to(device).contiguous()
        _weight_cache[key] = w
    return w


def image_to_patches(tensor: torch.Tensor, outer_patch_size: int, inner_patch_size: int):
    """tiling.py:10-64: ``[B,H,W]`` -> ``([B*nV*nH, outer, outer], [(nV, nH)] * B)``."""
    patches, grid_shape, _ = ops.image_to_patches(tensor.to(torch.float32).contiguous(), outer_patch_size,
                                                   inner_patch_size)
    return patches, [grid_shape] * tensor.shape[0]


def patches_to_image_weighted_average(tiles: torch.Tensor, image_information: Sequence[Tuple[int, int]],
                                      outer_patch_size: int, inner_patch_size: int, device=None) -> torch.Tensor:
    """tiling.py:91-140: weighted overlap average of the first image's patches -> ``[1, nV*inner, nH*inner]``."""
    nv, nh = image_information[0]
    tiles = tiles.to(torch.float32).contiguous()
    w = _weights_on(outer_patch_size, tiles.device)
    return ops.patches_to_image(tiles[: nv * nh], 1, (nv, nh), inner_patch_size, weights=w)


def patches_to_image(tiles: torch.Tensor, image_information: Sequence[Tuple[int, int]], outer_patch_size: int,
                     inner_patch_size: int) -> torch.Tensor:
    """tiling.py:143-181: unweighted overlap average -> ``[1, nV*inner, nH*inner]``."""
    nv, nh = image_information[0]
    tiles = tiles.to(torch.float32).contiguous()
    return ops.patches_to_image(tiles[: nv * nh], 1, (nv, nh), inner_patch_size)


def classify_patches(tile: torch.Tensor) -> int:
    """tiling.py:184-198: 0 if the tile is black (mean < 1e-10), else 1."""
    mask = ops.classify_patches(tile.to(torch.float32).contiguous().reshape(1, -1))
    return 0 if int(mask.item()) else 1


def filter_and_remember_black_patches(patches: torch.Tensor):
    """tiling.py:244-271: ``(non_black_patches, black_indices (list), original_shape)``.  One classifier launch
    instead of one device->host sync per patch."""
    mask = ops.classify_patches(patches.to(torch.float32).contiguous())
    keep = (mask == 0).nonzero(as_tuple=True)[0]
    black_indices: List[int] = mask.nonzero(as_tuple=True)[0].tolist()
    return patches.index_select(0, keep), black_indices, patches.shape


def reintegrate_black_patches(processed_patches: torch.Tensor, black_indices, original_shape) -> torch.Tensor:
    """tiling.py:274-303: zeros at ``black_indices``, processed patches elsewhere in order."""
    n = original_shape[0]
    full = torch.zeros((n, *processed_patches.shape[1:]), dtype=processed_patches.dtype,
                       device=processed_patches.device)
    keep = torch.ones(n, dtype=torch.bool, device=processed_patches.device)
    if len(black_indices):
        keep[torch.as_tensor(list(black_indices), dtype=torch.long, device=processed_patches.device)] = False
    full[keep] = processed_patches
    return full


def filter_black_patches_indices(undersampled: torch.Tensor) -> List[int]:
    """tiling.py:227-241: indices of the non-black patches (one classifier launch instead of a loop of syncs)."""
    mask = ops.classify_patches(undersampled.to(torch.float32).contiguous())
    return (mask == 0).nonzero(as_tuple=True)[0].tolist()


def filter_black_patches(undersampled: List[torch.Tensor], fullysampled: List[torch.Tensor]):
    """tiling.py:201-224: per image, drop the patches whose UNDERSAMPLED version is black from both lists."""
    for i in range(len(undersampled)):
        keep = (ops.classify_patches(undersampled[i].to(torch.float32).contiguous()) == 0).nonzero(as_tuple=True)[0]
        undersampled[i] = undersampled[i].index_select(0, keep)
        fullysampled[i] = fullysampled[i].index_select(0, keep.to(fullysampled[i].device))
    return undersampled, fullysampled


def extract_center_batch(batch: torch.Tensor, outer_patch_size: int, inner_patch_size: int) -> torch.Tensor:
    """tiling.py:306-322: the centre ``inner x inner`` window of every ``outer x outer`` patch (a view, as in the
    reference; used for the training target, src/train/training.py:190-196)."""
    padding = (outer_patch_size - inner_patch_size) // 2
    return batch[:, padding: padding + inner_patch_size, padding: padding + inner_patch_size]
