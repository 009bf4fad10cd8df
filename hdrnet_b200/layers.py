"""Layer helpers: drop-in for the slicing part of the reference's ``hdrnet/layers.py``.

Same names, argument order and shapes as the reference (hdrnet/layers.py:99-198) over
``torch.Tensor``:

    bilateral_slice(grid, guide)                                   layers.py:99-121
    bilateral_slice_apply(grid, guide, input_image, has_offset)    layers.py:125-148
    apply(sliced, input_image, has_affine_term)                    layers.py:153-198

``conv`` / ``fc`` (layers.py:25-93) live in ``hdrnet_b200.models`` next to the coefficient
network's kernels.
"""
from __future__ import annotations

import torch

from . import hdrnet_ops

__all__ = ["bilateral_slice", "bilateral_slice_apply", "apply"]


# pylint: disable=redefined-builtin
def bilateral_slice(grid: torch.Tensor, guide: torch.Tensor, name=None) -> torch.Tensor:
    """Slices into a bilateral grid using the guide map (hdrnet/layers.py:99-121).

    grid:  [B, gh, gw, gd, n_outputs] or the 6-D [B, gh, gw, gd, n_out, n_in];
    guide: [B, H, W].  Returns [B, H, W, n_outputs] or, for a 6-D grid, [B, H, W, n_out, n_in].
    A 6-D grid is packed input-channel-major, c = j * n_out + i, exactly as the reference's
    ``tf.concat(tf.unstack(grid, axis=5), 4)`` (layers.py:113-120).
    """
    del name
    six_d = grid.dim() == 6
    if six_d:
        B, gh, gw, gd, n_out, n_in = grid.shape
        grid = grid.permute(0, 1, 2, 3, 5, 4).reshape(B, gh, gw, gd, n_in * n_out)
    sliced = hdrnet_ops.bilateral_slice(grid, guide)
    if six_d:
        b, h, w, _ = sliced.shape
        sliced = sliced.reshape(b, h, w, n_in, n_out).permute(0, 1, 2, 4, 3)
    return sliced


def bilateral_slice_apply(grid: torch.Tensor, guide: torch.Tensor, input_image: torch.Tensor,
                          has_offset: bool = True, name=None) -> torch.Tensor:
    """Slices into a bilateral grid and applies the sliced affine model to ``input_image``
    in one fused pass (hdrnet/layers.py:125-148).

    grid: [B, gh, gw, gd, n_out * (n_in + has_offset)] or 6-D [B, gh, gw, gd, n_out, n_in(+1)]
    (flattened output-major, c = i * (n_in+1) + j, as ``tf.reshape`` does at layers.py:141-144);
    guide: [B, H, W]; input_image: [B, H, W, n_in].  Returns [B, H, W, n_out].
    """
    del name
    if grid.dim() == 6:
        B, gh, gw, gd, n_out, n_in = grid.shape
        grid = grid.reshape(B, gh, gw, gd, n_out * n_in)
    return hdrnet_ops.bilateral_slice_apply(grid, guide, input_image, has_offset=has_offset)


def apply(sliced: torch.Tensor, input_image: torch.Tensor, has_affine_term: bool = True,
          name=None) -> torch.Tensor:
    """Applies a sliced affine model to the input image (hdrnet/layers.py:153-198): the
    un-fused second half of bilateral_slice_apply, kept in plain torch ops as the reference
    keeps it in plain TF ops.

    sliced: [B, H, W, n_out, n_in(+1)]; input_image: [B, H, W, n_in].  Returns [B, H, W, n_out].
    """
    del name
    if input_image.dim() != 4:
        raise ValueError("input image should have dims [b,h,w,n_in].")
    in_shape = list(input_image.shape)
    sliced_shape = list(sliced.shape)
    if in_shape[:-1] != sliced_shape[:-2]:
        raise ValueError("input image and affine coefficients"
                         " dimensions do not match: {} and {}".format(in_shape, sliced_shape))
    n_in = sliced_shape[-1]
    if has_affine_term:
        n_in -= 1
    scale = sliced[..., :n_in]
    ret = (scale * input_image[:, :, :, None, :n_in]).sum(dim=-1)
    if has_affine_term:
        ret = ret + sliced[..., n_in]
    return ret
# pylint: enable=redefined-builtin
