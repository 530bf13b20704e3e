"""Grid-list <-> flat tensor helpers and shape checks.

Host-side mirror of the reference's `lightplane/misc_utils.py` (flatten_grid :25-46,
unflatten_grid :49-70, check_grid :106-140, check_grid_and_color_grid :143-198,
process_and_flatten_grid :201-234).

A *grid-list* is a list of 5-D tensors `[B, D_i, H_i, W_i, C]` (same B and C); its *flat* form
is one `[sum_i B*D_i*H_i*W_i, C]` row-major tensor plus a `[G, 5]` size table.  The CUDA
kernels only ever see the flat form ("Data layout in HBM").
"""

from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple

import torch


def assert_shape(x: torch.Tensor, shape: Tuple[int, ...]) -> None:
    assert tuple(x.shape) == tuple(shape), f"expected shape {tuple(shape)}, got {tuple(x.shape)}"


def if_not_none_else(x: Any, y: Any) -> Any:
    return y if x is None else x


def _sizes_as_lists(grid_sizes) -> List[List[int]]:
    """Normalise a size table given as tensor / list of lists / list of tuples to python ints.
    A CUDA tensor here costs one host sync; the functional ops avoid that by building the
    table from tensor shapes on the host."""
    if torch.is_tensor(grid_sizes):
        grid_sizes = grid_sizes.tolist()
    return [[int(v) for v in row] for row in grid_sizes]


def flatten_grid(grid: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Grid-list -> (flat `[sum BDHW, C]`, int32 sizes `[G,5]`) (misc_utils.py:25-46)."""
    dev = grid[0].device
    sizes = torch.tensor([list(g.shape) for g in grid], dtype=torch.int32, device=dev)
    chn = grid[0].shape[-1]
    if len(grid) == 1:
        flat = grid[0].reshape(-1, chn).contiguous()
    else:
        flat = torch.cat([g.reshape(-1, chn) for g in grid], dim=0).contiguous()
    return flat, sizes


def unflatten_grid(grid: torch.Tensor, grid_sizes) -> Tuple[torch.Tensor, ...]:
    """Inverse of `flatten_grid`; returns views into `grid` (misc_utils.py:49-70)."""
    sizes = _sizes_as_lists(grid_sizes)
    rows = [s[0] * s[1] * s[2] * s[3] for s in sizes]
    chunks = grid.split(rows, dim=0)
    return tuple(c.reshape(*s) for c, s in zip(chunks, sizes))


def pad_feature_to_block_size(feature: torch.Tensor, block_size: int) -> torch.Tensor:
    """Zero-pad dim 0 to a multiple of `block_size` (misc_utils.py:78-93)."""
    extra = (-feature.shape[0]) % int(block_size)
    if extra == 0:
        return feature
    tail = feature.new_zeros((extra,) + tuple(feature.shape[1:]))
    return torch.cat([feature, tail], dim=0)


def is_in_bounds(points: torch.Tensor) -> torch.Tensor:
    """`[..., 1]` bool: all coordinates inside the closed cube [-1, 1] (misc_utils.py:96-100)."""
    return (points.abs() <= 1.0).all(dim=-1, keepdim=True)


def _numel_of_sizes(grid_sizes) -> int:
    total = 0
    for s in _sizes_as_lists(grid_sizes):
        n = 1
        for v in s:
            n *= v
        total += n
    return total


def _check_list_against_sizes(grid: Sequence[torch.Tensor], grid_sizes) -> None:
    for g, s in zip(grid, _sizes_as_lists(grid_sizes)):
        assert_shape(g, tuple(s))


def check_grid(grid, grid_sizes=None):
    """Validate one grid given as list or flat tensor (misc_utils.py:106-140).  Like the
    reference only `list` (not `tuple`) is accepted as the list form -- plus tuples, which
    the reference rejects for no stated reason."""
    if isinstance(grid, (list, tuple)):
        grid = list(grid)
        if grid_sizes is not None:
            _check_list_against_sizes(grid, grid_sizes)
    elif torch.is_tensor(grid):
        assert grid_sizes is not None, "grid_sizes cannot be None when grid is a tensor"
        assert _numel_of_sizes(grid_sizes) == grid.numel(), (
            "grid_sizes has to be compatible to grid tensor shapes!"
        )
    else:
        raise NotImplementedError("grid should be either tensor or list")
    return grid, grid_sizes


def check_grid_and_color_grid(grid, color_grid, grid_sizes=None, color_grid_sizes=None):
    """Validate `grid` / `color_grid` pairs (misc_utils.py:143-198): same container type,
    same batch size and channel count, sizes consistent."""
    if isinstance(grid, tuple):
        grid = list(grid)
    if isinstance(color_grid, tuple):
        color_grid = list(color_grid)
    if color_grid is not None:
        assert type(grid) == type(color_grid), "grid and color_grid should have the same type"
    if isinstance(grid, list):
        if color_grid is not None:
            assert all(cg.shape[0] == g.shape[0] for cg, g in zip(color_grid, grid)), (
                "color_grid's batch size should be the same as grid's batch_size"
            )
            assert all(cg.shape[-1] == g.shape[-1] for cg, g in zip(color_grid, grid)), (
                "color_grid's feature dimension should be the same as grid's feature dimension"
            )
            if color_grid_sizes is not None:
                _check_list_against_sizes(color_grid, color_grid_sizes)
        if grid_sizes is not None:
            _check_list_against_sizes(grid, grid_sizes)
    elif torch.is_tensor(grid):
        check_grid(grid, grid_sizes)
        if color_grid is not None:
            assert color_grid_sizes is not None, (
                "color_grid_sizes cannot be None when color_grid is a tensor"
            )
            check_grid(color_grid, color_grid_sizes)
    else:
        raise NotImplementedError("grid should be either tensor or list")
    return grid, color_grid, grid_sizes, color_grid_sizes


def process_and_flatten_grid(grid, color_grid, grid_sizes=None, color_grid_sizes=None):
    """Bring grid (and colour grid) to the flat form + *host* size tables (python lists).

    Differs from the reference (misc_utils.py:201-234) in one deliberate way: the size table is
    returned as a python list of 5-int lists instead of a device tensor, so that the launch
    path never reads sizes back from the GPU.
    """
    if isinstance(grid, (list, tuple)):
        sizes = [list(g.shape) for g in grid]
        flat, _ = flatten_grid(list(grid))
        if color_grid is not None:
            csizes = [list(g.shape) for g in color_grid]
            cflat, _ = flatten_grid(list(color_grid))
        else:
            csizes, cflat = None, None
        return flat, cflat, sizes, csizes
    if torch.is_tensor(grid):
        sizes = _sizes_as_lists(grid_sizes)
        csizes = _sizes_as_lists(color_grid_sizes) if color_grid is not None else None
        flat = grid.reshape(-1, sizes[0][-1])
        cflat = color_grid.reshape(-1, csizes[0][-1]) if color_grid is not None else None
        return flat, cflat, sizes, csizes
    raise NotImplementedError("grid should be flatten either tensor or list")
