"""Direct C-ABI calls (include/lightplane_b200.h) on a golden case.  `lib` is either the product
library on CUDA tensors (-m gpu tests) or the test-only host emulation on CPU tensors."""
import torch

from lightplane_b200 import _cabi
from _golden import renderer_cfg, splat_cfg


def _dims(t):
    return [int(v) for v in t]


def decoder_spec(c):
    nh_t, nh_o, nh_c = _dims(c["n_hidden_trunk"]), _dims(c["n_hidden_opacity"]), _dims(c["n_hidden_color"])
    C = int(c["grid_sizes"][0][4])
    n_t, n_o, n_c = max(len(nh_t) - 1, 0), len(nh_o) - 1, len(nh_c) - 1
    hid_t = nh_t[1] if n_t else 0
    head_in = nh_o[0]
    return _cabi.DecoderSpec(
        n_t, n_o, n_c, hid_t, nh_o[1], nh_c[1], C if n_t else 0, head_in, nh_c[0],
        hid_t if n_t else 0, nh_c[-1], int(c["color_chn"]),
    )


def render_case(lib, c, device, ray_image_width=0):
    """forward + backward of a renderer golden case through the raw C-ABI."""
    f = lambda k: c[k].to(device=device, dtype=torch.float32).contiguous()
    cfgd = renderer_cfg(c)
    n = c["directions"].shape[0]
    color_chn = int(c["color_chn"])
    sizes = [_dims(s) for s in c["grid_sizes"]]
    dirs, orig, near, far, enc = f("directions"), f("origins"), f("near"), f("far"), f("encoding")
    gidx = c["grid_idx"].to(device=device, dtype=torch.int32).contiguous()
    grid, mlp = f("grid"), f("mlp_params")
    cgrid = f("color_grid") if "color_grid" in c else None
    scaf = f("scaffold") if "scaffold" in c else None
    cfg = _cabi.make_cfg(cfgd["num_samples"], cfgd["num_samples_inf"], cfgd["gain"],
                         cfgd["disparity_at_inf"], cfgd["mask_out_of_bounds_samples"],
                         cfgd["contract_coords"], cfgd["inject_noise_sigma"],
                         cfgd["inject_noise_seed"], n, ray_image_width)
    spec = decoder_spec(c)
    rays = _cabi.make_rays(dirs, orig, gidx, near, far, enc)
    gl = _cabi.make_grid_list(grid, sizes)
    cl = _cabi.make_grid_list(cgrid, sizes) if cgrid is not None else None
    sl = _cabi.make_grid_list(scaf, [list(scaf.shape) + [1]]) if scaf is not None else None
    B = _cabi.byref
    stream = _cabi.stream_ptr(torch.device(device))
    out_len = torch.empty(n, device=device)
    out_nlt = torch.empty(n, device=device)
    out_feat = torch.empty(n, color_chn, device=device)
    st = lib.lp_render_forward(stream, B(cfg), B(spec), B(rays), B(gl), B(cl), B(sl), mlp.data_ptr(),
                               out_len.data_ptr(), out_nlt.data_ptr(), out_feat.data_ptr(), color_chn)
    _cabi.check(lib, st, "lp_render_forward")
    g_grid, g_mlp, g_enc = torch.zeros_like(grid), torch.zeros_like(mlp), torch.empty_like(enc)
    g_cgrid = torch.zeros_like(cgrid) if cgrid is not None else None
    cl2, cf = f("cot_ray_length"), f("cot_features")
    cn = f("cot_nlt")
    st = lib.lp_render_backward(stream, B(cfg), B(spec), B(rays), B(gl), B(cl), B(sl), mlp.data_ptr(),
                                out_len.data_ptr(), out_feat.data_ptr(), color_chn, cl2.data_ptr(),
                                cn.data_ptr(), cf.data_ptr(), color_chn, g_grid.data_ptr(),
                                _cabi.ptr(g_cgrid), g_mlp.data_ptr(), g_enc.data_ptr())
    _cabi.check(lib, st, "lp_render_backward")
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    res = dict(ray_length=out_len, nlt=out_nlt, features=out_feat, g_grid=g_grid, g_mlp=g_mlp, g_enc=g_enc)
    if g_cgrid is not None:
        res["g_color_grid"] = g_cgrid
    return res


def splat_case(lib, c, device):
    f = lambda k: c[k].to(device=device, dtype=torch.float32).contiguous()
    kw = splat_cfg(c)
    n = c["directions"].shape[0]
    sizes = [_dims(s) for s in c["out_sizes"]]
    dirs, orig, near, far, feat = f("directions"), f("origins"), f("near"), f("far"), f("feature")
    gidx = c["grid_idx"].to(device=device, dtype=torch.int32).contiguous()
    rows = sum(s[0] * s[1] * s[2] * s[3] for s in sizes)
    C = sizes[0][4]
    out = torch.zeros(rows, C, device=device)
    wgt = torch.zeros(rows, 1, device=device)
    cfg = _cabi.make_cfg(kw["num_samples"], kw["num_samples_inf"], 1.0, 1e-5,
                         kw["mask_out_of_bounds_samples"], kw["contract_coords"], 0.0, 0, n)
    rays = _cabi.make_rays(dirs, orig, gidx, near, far, feat)
    ol = _cabi.make_grid_list(out, sizes)
    B = _cabi.byref
    stream = _cabi.stream_ptr(torch.device(device))
    use_mlp = "mlp_params" in c
    if use_mlp:
        nh = _dims(c["n_hidden"])
        spec = _cabi.MlpSpec(len(nh) - 1, nh[0], nh[1] if len(nh) > 2 else nh[-1], nh[-1])
        mlp, ing = f("mlp_params"), f("input_grid")
        in_sizes = [_dims(s) for s in c["input_sizes"]]
        il = _cabi.make_grid_list(ing, in_sizes)
        st = lib.lp_mlp_splat_forward(stream, B(cfg), B(spec), B(rays), None, B(il), mlp.data_ptr(),
                                      B(ol), wgt.data_ptr())
        _cabi.check(lib, st, "lp_mlp_splat_forward")
    else:
        st = lib.lp_splat_forward(stream, B(cfg), B(rays), None, B(ol), wgt.data_ptr())
        _cabi.check(lib, st, "lp_splat_forward")
    st = lib.lp_splat_normalize(stream, out.data_ptr(), wgt.data_ptr(), rows, C)
    _cabi.check(lib, st, "lp_splat_normalize")
    g = (f("cot") / wgt).contiguous()
    gl = _cabi.make_grid_list(g, sizes)
    g_feat = torch.empty_like(feat)
    res = dict(out=out)
    if use_mlp:
        g_mlp, g_in = torch.zeros_like(mlp), torch.zeros_like(ing)
        st = lib.lp_mlp_splat_backward(stream, B(cfg), B(spec), B(rays), None, B(il), mlp.data_ptr(),
                                       B(gl), g_feat.data_ptr(), g_mlp.data_ptr(), g_in.data_ptr())
        _cabi.check(lib, st, "lp_mlp_splat_backward")
        res.update(g_mlp=g_mlp, g_input_grid=g_in)
    else:
        st = lib.lp_splat_backward(stream, B(cfg), B(rays), None, B(gl), g_feat.data_ptr())
        _cabi.check(lib, st, "lp_splat_backward")
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    res["g_feat"] = g_feat
    return res
