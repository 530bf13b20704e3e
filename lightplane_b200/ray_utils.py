"""Ray containers and ray-side helpers of the Renderer / Splatter hot path.

Host-side mirror of the reference's `lightplane/ray_utils.py` (class `Rays` :19-178,
`calc_harmonic_embedding` :181-212, `jitter_near_far` :220-229).  Same names, fields and
argument meaning, so a `Rays` built for the reference works here unchanged; the functional
ops only duck-type on the six field names.
"""

from __future__ import annotations

import copy
import dataclasses
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

_TENSOR_FIELDS = ("directions", "origins", "grid_idx", "near", "far", "encoding")


@dataclass
class Rays:
    """A batch of `N` rays `x(t) = origin + t * direction`, `t in [near, far]`.

    directions [N,3], origins [N,3] (float), grid_idx [N] (integer scene index into the
    grid batch), near/far [N], encoding [N,E] or None.  Directions need not be normalised.
    (reference: ray_utils.py:19-57)
    """

    directions: torch.Tensor
    origins: torch.Tensor
    grid_idx: torch.Tensor
    near: torch.Tensor
    far: torch.Tensor
    encoding: Optional[torch.Tensor] = None

    def __post_init__(self):
        _check_ray_fields(self)

    # ---- introspection -------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.directions.device

    @property
    def num_rays(self) -> int:
        return int(self.directions.shape[0])

    def _items(self):
        for name in _TENSOR_FIELDS:
            yield name, getattr(self, name)

    # ---- container protocol --------------------------------------------------------
    def __getitem__(self, key) -> "Rays":
        """Sub-select rays (reference: ray_utils.py:90-107)."""
        picked = {k: (None if v is None else v[key]) for k, v in self._items()}
        return type(self)(**picked)

    def pad_to_block_size(self, block_size: int) -> Tuple["Rays", int]:
        """Zero-pad every field so that N is a multiple of `block_size`.

        Returns the (possibly new) rays and the number of rays that were appended
        (reference: ray_utils.py:109-140).  The CUDA path itself accepts any N; this is
        kept for API parity and for the noise index convention, which counts padded rays.
        """
        n = self.num_rays
        extra = (-n) % int(block_size)
        if extra == 0:
            return self, 0
        padded = {}
        for k, v in self._items():
            if v is None:
                padded[k] = None
            else:
                tail = v.new_zeros((extra,) + tuple(v.shape[1:]))
                padded[k] = torch.cat([v, tail], dim=0)
        return type(self)(**padded), extra

    def to(self, device, copy: bool = False) -> "Rays":
        """Move all fields to `device` (the reference's version :142-169 only handles the
        same-device early return; this one performs the move)."""
        device = torch.device(device)
        if not copy and self.device == device:
            return self
        moved = {
            k: (None if v is None else v.to(device=device, copy=copy))
            for k, v in self._items()
        }
        return type(self)(**moved)

    def clone(self) -> "Rays":
        return copy.deepcopy(self)


def _check_ray_fields(r: "Rays") -> None:
    """Shape / dtype / device validation (reference: ray_utils.py:232-274)."""
    d, o, gi, nr, fr, enc = r.directions, r.origins, r.grid_idx, r.near, r.far, r.encoding
    assert d.ndim == 2 and o.ndim == 2, "directions / origins must be [N, 3]"
    assert d.shape[1] == 3 and o.shape[1] == 3, "directions / origins must be [N, 3]"
    assert gi.ndim == 1 and nr.ndim == 1 and fr.ndim == 1, "grid_idx / near / far must be [N]"
    assert not gi.is_floating_point(), "grid_idx must be an integer tensor"
    n, dev = d.shape[0], d.device
    for name, v in (("origins", o), ("grid_idx", gi), ("near", nr), ("far", fr)):
        assert v.device == dev, f"{name} is on a wrong device ({v.device}, expected {dev})"
        assert v.shape[0] == n, f"Unexpected number of elements in {name} ({v.shape[0]}, expected {n})"
    if enc is not None:
        assert enc.ndim == 2 and enc.shape[0] == n, "encoding must be [N, E]"
        assert enc.device == dev, "encoding is on a wrong device"


def calc_harmonic_embedding_dim(n_harmonic_functions: int) -> int:
    """3 raw coordinates + (sin, cos) x 3 coordinates x n frequencies (ray_utils.py:215-217)."""
    return 3 + 6 * int(n_harmonic_functions)


def calc_harmonic_embedding(directions: torch.Tensor, n_harmonic_functions: int) -> torch.Tensor:
    """NeRF-style harmonic embedding `[sin(2^k d), cos(2^k d), d]` (ray_utils.py:181-212).

    Output layout (last dim): for phase in (0, pi/2): for coordinate in xyz: for k in
    0..n-1 -> sin(d_coord * 2^k + phase); then the 3 raw coordinates.
    """
    n = int(n_harmonic_functions)
    if n == 0:
        return directions
    freqs = torch.pow(
        torch.tensor(2.0, dtype=directions.dtype, device=directions.device),
        torch.arange(n, dtype=directions.dtype, device=directions.device),
    )
    scaled = directions.unsqueeze(-1) * freqs  # [..., 3, n]
    phases = torch.tensor([0.0, 0.5 * math.pi], dtype=directions.dtype, device=directions.device)
    waves = torch.sin(scaled.unsqueeze(-3) + phases.view(2, 1, 1))  # [..., 2, 3, n]
    waves = waves.reshape(*directions.shape[:-1], 6 * n)
    return torch.cat([waves, directions], dim=-1)


def jitter_near_far(near: torch.Tensor, far: torch.Tensor, num_samples: int):
    """Shift near and far by the same uniform offset in +-(far-near)/num_samples
    (ray_utils.py:220-229)."""
    step = (far - near) / num_samples
    shift = (torch.rand_like(near) * 2.0 - 1.0) * step
    return near + shift, far + shift
