"""lightplane_b200 -- B200-native drop-in for the Lightplane Renderer / Splatter hot path.

Public surface = the reference's (`lightplane/__init__.py:8-31`) minus the naive
implementations (they live in `oracle/` as test infrastructure) and the plotly visualiser
(out of scope).  Importing the package does not load the CUDA library; the first op call does
and raises if it is missing -- there is no CPU fallback.
"""

from .lightplane_renderer import LightplaneFunction, lightplane_renderer
from .lightplane_splatter import (
    LightplaneSplatterFunction,
    lightplane_mlp_splatter,
    lightplane_splatter,
)
from .misc_utils import flatten_grid, unflatten_grid
from .mlp_utils import (
    DecoderParams,
    SplatterParams,
    flatten_decoder_params,
    flatten_splatter_params,
    flattened_decoder_params_to_list,
    flattened_triton_decoder_to_list,
    get_triton_function_input_dims,
    init_decoder_params,
    init_splatter_params,
)
from .ray_utils import Rays, calc_harmonic_embedding, calc_harmonic_embedding_dim, jitter_near_far
from .renderer_module import LightplaneRenderer
from .splatter_module import LightplaneMLPSplatter, LightplaneSplatter

__version__ = "0.1.0"

__all__ = [
    "DecoderParams",
    "LightplaneFunction",
    "LightplaneMLPSplatter",
    "LightplaneRenderer",
    "LightplaneSplatter",
    "LightplaneSplatterFunction",
    "Rays",
    "SplatterParams",
    "calc_harmonic_embedding",
    "calc_harmonic_embedding_dim",
    "flatten_decoder_params",
    "flatten_grid",
    "flatten_splatter_params",
    "flattened_decoder_params_to_list",
    "flattened_triton_decoder_to_list",
    "get_triton_function_input_dims",
    "init_decoder_params",
    "init_splatter_params",
    "jitter_near_far",
    "lightplane_mlp_splatter",
    "lightplane_renderer",
    "lightplane_splatter",
    "unflatten_grid",
]
