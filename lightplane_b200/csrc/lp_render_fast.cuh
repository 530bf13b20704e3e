// Specialised tensor-core Renderer kernels (default decoder shape).  Placeholder until the fast
// path lands: nothing is claimed as supported, so every call takes the generic kernels.
#pragma once
#include "lp_render_generic.cuh"

static inline bool lp_fast_render_supported(const LpRenderArgs&) { return false; }
static inline int lp_fast_render_forward(cudaStream_t, const LpRenderArgs&, const float*, float*, float*, float*, int) { return LP_ERR_UNSUPPORTED; }
static inline int lp_fast_render_backward(cudaStream_t, const LpRenderArgs&, const float*, const LpBwdIo&) { return LP_ERR_UNSUPPORTED; }
