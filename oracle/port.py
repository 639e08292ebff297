"""CPU oracle: a functional restatement of the 3DHumanGAN generator/discriminator hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; the product path
(`3dhumangan_b200/`) never does and fails loudly without its CUDA library.

Every function restates the arithmetic of one reference function (file:line under the
reference checkout) as pure functions over a flat `params` dict keyed by the reference's own
`state_dict()` names.  Random draws (ray jitter, sigma noise) are INPUTS, so the oracle and the
CUDA path can be fed identical values.

Pinning status ("Oracle"):
  * pinned against the unmodified reference modules imported in the build container
    (`tests/test_oracle_pin.py`, fixtures from `tests/golden/make_golden.py`);
  * the K=1 nearest-vertex search is `pytorch3d.ops.knn_points` (pytorch3d 0.6.2, not vendored by
    the reference, no reference test at that boundary) => that sub-step is "parity unpinned":
    it follows pytorch3d's documented contract (squared distances by direct differences in fp32,
    lowest index on ties).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LRELU = 0.2


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def normalize_2nd_moment(x, dim=1, eps=1e-8):
    """lib/components/util.py:58-59."""
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def bias_act_ref(x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None):
    """lib/components/ops/bias_act.py:90-121 (`_bias_act_ref`) for the activations it defines."""
    table = {
        "linear": (lambda v, a: v, 1.0, 0.0),
        "relu": (lambda v, a: torch.relu(v), math.sqrt(2), 0.0),
        "lrelu": (lambda v, a: F.leaky_relu(v, a), math.sqrt(2), 0.2),
        "tanh": (lambda v, a: torch.tanh(v), 1.0, 0.0),
        "sigmoid": (lambda v, a: torch.sigmoid(v), 1.0, 0.0),
        "elu": (lambda v, a: F.elu(v), 1.0, 0.0),
        "selu": (lambda v, a: F.selu(v), 1.0, 0.0),
        "softplus": (lambda v, a: F.softplus(v), 1.0, 0.0),
        "swish": (lambda v, a: torch.sigmoid(v) * v, math.sqrt(2), 0.0),
    }
    fn, def_gain, def_alpha = table[act]
    alpha = float(alpha if alpha is not None else def_alpha)
    gain = float(gain if gain is not None else def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        shape = [-1 if i == dim else 1 for i in range(x.ndim)]
        x = x + b.reshape(shape)
    x = fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


# --------------------------------------------------------------------------------------------
# mapping networks  (a1, a2 in SURVEY.md §8a)
# --------------------------------------------------------------------------------------------
def mapping_network(params, z, prefix="neural_field_mapping_network."):
    """lib/components/mapping_networks.py:33-41.  z [B,L] -> freq, phase [B, 4H] each."""
    x = normalize_2nd_moment(z.to(torch.float32))
    for i in (0, 2, 4):
        x = F.leaky_relu(F.linear(x, params[f"{prefix}network.{i}.weight"], params[f"{prefix}network.{i}.bias"]), LRELU)
    fp = F.linear(x, params[f"{prefix}network.6.weight"], params[f"{prefix}network.6.bias"])
    half = fp.shape[-1] // 2
    return fp[..., :half], fp[..., half:]


def _fc(params, name, x, act, lr_mul=0.01, extra_gain=1.0):
    """lib/components/mapping_networks.py:92-121 (`FullyConnectedLayer.forward`)."""
    w = params[name + ".weight"]
    b = params[name + ".bias"]
    w = w * (lr_mul / math.sqrt(w.shape[1]) * extra_gain)
    b = b * lr_mul
    if act == "linear":
        return torch.addmm(b[None], x, w.t())
    return bias_act_ref(x.matmul(w.t()), b, act=act)


def synthesis_mapping(params, z, prefix="synthesis_mapping_network.", trunk_layers=7):
    """lib/components/mapping_networks.py:184-216 with branch_layers=1, num_ws=1, c_dim=0.

    Returns the 'superres' styles [B,1,F]; the 'implicit' branch output is discarded by the
    generator (map3d_generator.py:222)."""
    x = normalize_2nd_moment(z.to(torch.float32))
    for i in range(trunk_layers):
        x = _fc(params, f"{prefix}trunk{i}", x, "lrelu")
    s = _fc(params, f"{prefix}superres0", x, "lrelu")
    return s[:, None, :]


# --------------------------------------------------------------------------------------------
# rays, jitter, camera transform  (a3, a4, a5)
# --------------------------------------------------------------------------------------------
def initial_rays(focals, scales, num_steps, render_width, render_height, ray_start, ray_end):
    """lib/generators/volume_rendering.py:86-110.  Ray r = h*Rw + w."""
    B = focals.shape[0]
    W, H = render_width, render_height
    dev = focals.device                                   # device-aware: the GPU tests run this checker on the device
    xs = torch.linspace(-W / H, W / H, W, device=dev)
    ys = torch.linspace(-1, 1, H, device=dev)
    x = xs[None, :].expand(H, W).reshape(-1)
    y = ys[:, None].expand(H, W).reshape(-1)
    xyz = torch.stack([x[None].expand(B, -1), y[None].expand(B, -1), focals[:, None].expand(B, H * W)], -1)
    d = xyz / (torch.norm(xyz, dim=-1, keepdim=True) + 1e-12)          # util.py:87-91
    z = torch.linspace(ray_start, ray_end, num_steps, device=dev).reshape(1, 1, num_steps, 1)
    z = z.expand(B, H * W, num_steps, 1) + (focals / scales).view(B, 1, 1, 1)
    pts = d[:, :, None, :] * z
    return pts, z, d


def jitter_and_transform(pts, z, d, cam2world, u):
    """volume_rendering.py:124-130 (perturb, `u` = the torch.rand draw) and :150-155 (cam->world)."""
    delta = z[:, :, 1:2, :] - z[:, :, 0:1, :]
    off = (u - 0.5) * delta
    z = z + off
    pts = pts + off * d[:, :, None, :]
    B, R, S, _ = pts.shape
    hom = F.pad(pts, (0, 1), value=1.0).reshape(B, -1, 4).permute(0, 2, 1)
    w = torch.bmm(cam2world, hom).permute(0, 2, 1).reshape(B, R, S, 4)[..., :3]
    return w, z


# --------------------------------------------------------------------------------------------
# geometry features  (a6)
# --------------------------------------------------------------------------------------------
def knn1(points, vertices, chunk=4096):
    """K=1 nearest vertex: contract of pytorch3d.ops.knn_points as called at lib/components/smpl.py:220.

    d2 = (dx*dx + dy*dy) + dz*dz with one fp32 rounding per operation; lowest index on ties."""
    B, N, _ = points.shape
    d2min = torch.empty(B, N, dtype=torch.float32, device=points.device)
    idx = torch.empty(B, N, dtype=torch.int64, device=points.device)
    for b in range(B):
        vx, vy, vz = (vertices[b, :, k][None].float() for k in range(3))
        for s in range(0, N, chunk):
            q = points[b, s:s + chunk].float()
            dx, dy, dz = q[:, 0:1] - vx, q[:, 1:2] - vy, q[:, 2:3] - vz
            d2 = dx * dx + dy * dy + dz * dz
            m = d2.min(dim=1).values
            d2min[b, s:s + chunk] = m
            idx[b, s:s + chunk] = (d2 == m[:, None]).to(torch.int8).argmax(dim=1)
    return d2min, idx


def geo_features(points, skeletons, vertices, tpose_vertices, fk_matrices, lbs_weights, legacy_mode=False):
    """lib/components/smpl.py:210-249.  Returns ([B,N,31] features, [B,N] int64 nearest index)."""
    B, N, _ = points.shape
    V = vertices.shape[1]
    joint_d = torch.cdist(points, skeletons) / 2.4
    ik = torch.inverse(fk_matrices.float())
    vertex_ik = torch.einsum("bij,bjkl->bikl", lbs_weights, ik).reshape(B, V, 16)
    d2, idx = knn1(points, vertices)
    pik = torch.gather(vertex_ik, 1, idx[:, :, None].expand(B, N, 16)).reshape(B, N, 4, 4)
    hom = F.pad(points, (0, 1), value=1.0)
    cano = torch.einsum("bijk,bik->bij", pik, hom)[..., :3].clone()
    cano[..., 0] = cano[..., 0] / 2.0
    cano[..., 1] = (cano[..., 1] + 0.2) / 2.0
    cano[..., 2] = cano[..., 2] / 1.3
    cv = torch.gather(tpose_vertices, 1, idx[:, :, None].expand(B, N, 3)).clone()
    cv[..., 2] = cv[..., 2] / 0.2
    nd = torch.sqrt(d2)[..., None] / 1.3
    parts = [joint_d, cano, cv, nd] if legacy_mode else [cano, joint_d, cv, nd]
    return torch.cat(parts, -1), idx


# --------------------------------------------------------------------------------------------
# per-point FiLM-SIREN  (a7)
# --------------------------------------------------------------------------------------------
def siren(params, pts, freq, phase, geo, dirs, input_scaler, hidden_dim, num_blocks=4, prefix="neural_field."):
    """lib/implicit_funcitions/modulated.py:41-75 with pigan_layers.py:63-87.

    pts [B,N,3], geo [B,N,G], dirs [B,N,3], freq/phase [B, num_blocks*H] -> [B,N,3+F+1] = rgb, feat, sigma."""
    P = lambda n: params[prefix + n]
    H = hidden_dim
    f = freq * 15 + 30
    a = torch.sin(30.0 * F.linear(pts * input_scaler, P("first_layer_coord.layer.weight"), P("first_layer_coord.layer.bias")))
    g = torch.sin(30.0 * F.linear(geo, P("first_layer_mod.layer.weight"), P("first_layer_mod.layer.bias")))
    x = torch.cat([a, g], -1)
    for i in range(num_blocks):
        y = F.linear(x, P(f"network.{i}.layer.weight"), P(f"network.{i}.layer.bias"))
        x = torch.sin(f[:, None, i * H:(i + 1) * H] * y + phase[:, None, i * H:(i + 1) * H])
    sigma = F.linear(x, P("sigma_layer.weight"), P("sigma_layer.bias"))
    y = F.linear(torch.cat([dirs, x], -1), P("color_layer_sine.layer.weight"), P("color_layer_sine.layer.bias"))
    c = torch.sin(f[:, None, -H:] * y + phase[:, None, -H:])          # re-uses the last slice (:68)
    rgb = torch.sigmoid(F.linear(c, P("color_layer_linear.weight"), P("color_layer_linear.bias")))
    feat = F.linear(c, P("feature_layer_linear.weight"), P("feature_layer_linear.bias"))
    return torch.cat([rgb, feat, sigma], -1)


# --------------------------------------------------------------------------------------------
# volume integration  (a8)
# --------------------------------------------------------------------------------------------
def ray_integration(out, z, noise, noise_std, white_back, last_back, clamp_mode="relu"):
    """lib/generators/volume_rendering.py:12-56.  out [B,R,S,C+1], z [B,R,S,1], noise [B,R,S,1] ~ N(0,1)."""
    feats, sig = out[..., :-1], out[..., -1:]
    delta = z[:, :, 1:] - z[:, :, :-1]
    delta = torch.cat([delta, 1e9 * torch.ones_like(delta[:, :, :1])], -2)
    pre = sig + noise * noise_std
    dens = F.relu(pre) if clamp_mode == "relu" else F.softplus(pre)
    alpha = 1 - torch.exp(-delta * dens)
    shifted = torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-12], -2)
    w = alpha * torch.cumprod(shifted, -2)[:, :, :-1]
    wsum = w.sum(2)
    if last_back:
        w = w.clone()
        w[:, :, -1] += 1 - wsum
        rgbf = (w * feats).sum(-2)
        depth = (w * z).sum(-2)
    else:
        rgbf = (w * feats).sum(-2)
        wd = w.clone()
        wd[:, :, -1] += 1 - wsum
        depth = (wd * z).sum(-2)
    if white_back:
        rgbf = rgbf + 1 - wsum
    return rgbf, depth, w


def render(params, freq, phase, cond, cfg, u, noise, dtype=None):
    """lib/generators/map3d_generator.py:381-523 with hierarchical_sample=False, staged=False.

    `u` [B,R,S,1] is the jitter draw, `noise` [B,R,S,1] the sigma-noise draw.
    Returns rgb_render [B,3,Rh,Rw], feature_maps [B,F,Rh,Rw], depth [B,R,1], weights, nearest idx.
    `dtype` (tests only): evaluate the MLP and the integration in that dtype (parameters given in it); rays and geometry
    features stay fp32, as the reference computes them (`.float()` casts at smpl.py:217,220)."""
    Rw, Rh, S = cfg["render_width"], cfg["render_height"], cfg["num_steps"]
    Fd, H = cfg["feature_dim"], cfg["hidden_dim"]
    focals = cond["intrinsics"][:, 0, 0]
    scales = cond["scales"].float()
    B = freq.shape[0]
    pts, z, d = initial_rays(focals, scales, S, Rw, Rh, cfg["ray_start"], cfg["ray_end"])
    pw, z = jitter_and_transform(pts, z, d, cond["cam2world_matrices"], u)
    pw = pw.reshape(B, Rw * Rh * S, 3)
    if cfg.get("lock_view_dependence", False):
        dirs = torch.zeros_like(pw)
        dirs[..., -1] = -1
    else:
        dw = torch.bmm(cond["cam2world_matrices"][:, :3, :3], d.permute(0, 2, 1)).permute(0, 2, 1)
        dirs = dw[:, :, None, :].expand(B, Rw * Rh, S, 3).reshape(B, -1, 3)
    geo, idx = geo_features(pw, cond["skeletons_xyz"], cond["vertices"], cond["tpose_vertices"],
                            cond["fk_matrices"], cond["lbs_weights"], cfg.get("legacy_mode", False))
    if dtype is not None:
        pw, geo, dirs, z, noise, freq, phase = (t.to(dtype) for t in (pw, geo, dirs, z, noise, freq, phase))
    out = siren(params, pw, freq, phase, geo, dirs, 2.0 / cfg["side_length"], H, cfg["neural_field_blocks"])
    out = out.reshape(B, Rw * Rh, S, Fd + 4)
    rgbf, depth, w = ray_integration(out, z, noise, cfg["nerf_noise"], cfg.get("white_back", False),
                                     cfg.get("last_back", False), cfg["clamp_mode"])
    img = rgbf.reshape(B, Rh, Rw, Fd + 3).permute(0, 3, 1, 2)
    return img[:, :3] * 2 - 1, img[:, 3:], depth, w, idx


# --------------------------------------------------------------------------------------------
# synthesis network  (a10 - a13, a15)
# --------------------------------------------------------------------------------------------
def spectral_weight(w_orig, u, v, training, eps=1e-12):
    """torch.nn.utils.spectral_norm (one power iteration per training forward), as applied at
    lib/components/map3d_layers.py:205-206 and lib/discriminators/unet_discriminators.py:18.
    Returns (W / sigma, u', v')."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():      # the power iteration is not differentiated (spectral_norm.py: `with torch.no_grad()`)
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=eps)
    sigma = torch.dot(u, torch.mv(wm, v))
    return w_orig / sigma, u, v


def batch_norm_2d(x, weight, bias, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    """nn.SyncBatchNorm in a single process == batch statistics over (B,H,W) (map3d_layers.py:162).
    Returns (y, new_running_mean, new_running_var)."""
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() / x.shape[1]
        rm = (1 - momentum) * running_mean + momentum * mean
        rv = (1 - momentum) * running_var + momentum * var * n / (n - 1)
    else:
        mean, var, rm, rv = running_mean, running_var, running_mean, running_var
    y = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + eps)
    return y * weight[None, :, None, None] + bias[None, :, None, None], rm, rv


def spade_half(params, pfx_spade, pfx_conv, x, style, training, stats_out=None):
    """One SPADE2d (map3d_layers.py:176-190) + LeakyReLU(0.2) + spectral-normed 1x1 conv (:226-232)."""
    P = lambda n: params[n]
    xn, rm, rv = batch_norm_2d(x, P(pfx_spade + "first_norm.weight"), P(pfx_spade + "first_norm.bias"),
                               P(pfx_spade + "first_norm.running_mean"), P(pfx_spade + "first_norm.running_var"), training)
    if stats_out is not None:
        stats_out[pfx_spade + "first_norm.running_mean"] = rm
        stats_out[pfx_spade + "first_norm.running_var"] = rv
    actv = F.relu(F.conv2d(style, P(pfx_spade + "mlp_shared.0.weight"), P(pfx_spade + "mlp_shared.0.bias")))
    gamma = 1 + F.conv2d(actv, P(pfx_spade + "mlp_gamma.weight"), P(pfx_spade + "mlp_gamma.bias"))
    beta = F.conv2d(actv, P(pfx_spade + "mlp_beta.weight"), P(pfx_spade + "mlp_beta.bias"))
    y = F.leaky_relu(xn * gamma + beta, LRELU)
    w, u, v = spectral_weight(P(pfx_conv + "weight_orig"), P(pfx_conv + "weight_u"), P(pfx_conv + "weight_v"), training)
    if stats_out is not None:
        stats_out[pfx_conv + "weight_u"] = u
        stats_out[pfx_conv + "weight_v"] = v
    return F.conv2d(y, w, P(pfx_conv + "bias"))


def synthesis_network(params, x, style, fixed_style, cfg, training=True, stats_out=None,
                      prefix="synthesis_network.", return_internal=False):
    """lib/generators/map3d_generator.py:58-97 + SPADEBlock.forward (map3d_layers.py:218-238) + ToRGB (:346-352)."""
    nb = cfg["synthesis_blocks"]
    mode = cfg.get("map3d_mode", "isolated")
    mod_blocks = cfg["mod_blocks"]
    B, C = fixed_style.shape[0], fixed_style.shape[2]
    fs_map = fixed_style.view(B, C, 1, 1).expand_as(style)
    rgb = None
    internal = {}
    for k in range(nb):
        if mode == "all":
            s = style + fs_map
        elif mode == "mixed":
            s = (style if k in mod_blocks else torch.zeros_like(style)) + fs_map
        elif mode == "isolated":
            s = style if k in mod_blocks else fs_map
        else:
            raise ValueError("invalid map3d_mode")
        blk = f"{prefix}network.m3d_{k}."
        x_in = x
        x = spade_half(params, blk + "spade_0.", blk + "conv_0.", x, s, training, stats_out)
        x = spade_half(params, blk + "spade_1.", blk + "conv_1.", x, s, training, stats_out)
        if k >= nb // 2 and x.shape[-1] == x_in.shape[-1]:
            x = x + x_in
        if k >= nb // 2 - 1:
            t = F.conv2d(x, params[f"{prefix}to_rgbs.m3d_{k}.linear.weight"], params[f"{prefix}to_rgbs.m3d_{k}.linear.bias"])
            rgb = t if rgb is None else t + rgb
        if return_internal:
            internal[f"m3d_{k}"] = x
    return (rgb, internal) if return_internal else rgb


def synthesis_input(params, B, Hg, Wg, prefix="synthesis_input."):
    """SynthesisInput.get_2d_coords + forward (map3d_layers.py:260-275): sin(Conv1x1_{2->F}(coords))."""
    w = params[prefix + "network.0.weight"]
    i = torch.linspace(-1, 1, Hg, device=w.device, dtype=torch.float32).to(w.dtype)
    j = torch.linspace(-1, 1, Wg, device=w.device, dtype=torch.float32).to(w.dtype)
    coords = torch.stack([i[:, None].expand(Hg, Wg), j[None, :].expand(Hg, Wg)], 0)[None].repeat(B, 1, 1, 1)
    return torch.sin(F.conv2d(coords, params[prefix + "network.0.weight"], params[prefix + "network.0.bias"]))


def generator_forward(params, z, cond, cfg, u, noise, training=True, stats_out=None, dtype=None):
    """Map3DGenerator.forward (map3d_generator.py:208-280), render + synthesis path.

    Returns dict(rgbs, rgbs_render, feature_maps, depths, nearest_idx).  `dtype`: see `render` (tests: an fp64 control)."""
    zz = z if cfg.get("neural_field_latent_input", True) else torch.zeros_like(z)
    freq, phase = mapping_network(params, zz)
    styles = synthesis_mapping(params, z)
    if dtype is not None:
        styles = styles.to(dtype)
    rgb_r, fmap, depth, w, idx = render(params, freq, phase, cond, cfg, u, noise, dtype=dtype)
    Hg, Wg = cfg["gen_height"], cfg["gen_width"]
    style = F.interpolate(fmap, (Hg, Wg), mode="bilinear")             # :244-245 (align_corners=False)
    x0 = synthesis_input(params, z.shape[0], Hg, Wg)
    rgb = synthesis_network(params, x0, style, styles, cfg, training, stats_out)
    return {"rgbs": rgb, "rgbs_render": rgb_r, "feature_maps": fmap, "depths": depth, "nearest_idx": idx}


# --------------------------------------------------------------------------------------------
# discriminator  (a14)
# --------------------------------------------------------------------------------------------
def _sn_conv(params, name, x, training, padding, stats_out=None, spectral=True):
    if spectral:
        w, u, v = spectral_weight(params[name + ".weight_orig"], params[name + ".weight_u"], params[name + ".weight_v"], training)
        if stats_out is not None:
            stats_out[name + ".weight_u"], stats_out[name + ".weight_v"] = u, v
    else:
        w = params[name + ".weight"]
    return F.conv2d(x, w, params[name + ".bias"], padding=padding)


def res_block(params, name, x, up_or_down, first, learned_shortcut, training, stats_out=None):
    """lib/discriminators/unet_discriminators.py:47-71."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    pool = lambda t: F.avg_pool2d(t, 2)
    # shortcut (:57-71)
    s = x
    if first:
        if up_or_down < 0:
            s = pool(s)
        if learned_shortcut:
            s = _sn_conv(params, name + ".conv_s", s, training, 0, stats_out)
    else:
        if up_or_down > 0:
            s = up(s)
        if learned_shortcut:
            s = _sn_conv(params, name + ".conv_s", s, training, 0, stats_out)
        if up_or_down < 0:
            s = pool(s)
    # residual (:20-38)
    if first:
        dx = _sn_conv(params, name + ".conv1", x, training, 1, stats_out)
    elif up_or_down > 0:
        dx = _sn_conv(params, name + ".conv1.2", up(F.leaky_relu(x, LRELU)), training, 1, stats_out)
    else:
        dx = _sn_conv(params, name + ".conv1.1", F.leaky_relu(x, LRELU), training, 1, stats_out)
    dx = _sn_conv(params, name + ".conv2.1", F.leaky_relu(dx, LRELU), training, 1, stats_out)
    if up_or_down < 0:
        dx = pool(dx)
    return s + dx


def discriminator_forward(params, images, cfg, training=True, stats_out=None):
    """UNetDiscriminator.forward (unet_discriminators.py:125-160)."""
    nb = min(cfg.get("discriminator_blocks", 6), int(math.log2(max(cfg["gen_height"], cfg["gen_width"]))) - 1)
    ch = [3, 128, 128, 256, 256, 512, 512, 512, 512]
    x = images
    skips = []
    for i in range(nb):
        x = res_block(params, f"body_down.{i}", x, -1, i == 0, ch[i] != ch[i + 1], training, stats_out)
        skips.append(x)
    if min(x.shape[2:4]) > 1:
        latents = F.conv2d(x, params["latent_layer.weight"], params["latent_layer.bias"]).view(x.shape[0], -1)
    else:
        latents = torch.zeros(x.shape[0], cfg["latent_dim"], dtype=x.dtype, device=x.device)
    outs = [ch[nb - 1]] + [ch[nb - i - 1] for i in range(1, nb - 1)] + [64]
    ins = [ch[nb]] + [2 * ch[nb - i] for i in range(1, nb - 1)] + [2 * ch[1]]
    x = res_block(params, "body_up.0", x, 1, False, ins[0] != outs[0], training, stats_out)
    for i in range(1, nb):
        x = res_block(params, f"body_up.{i}", torch.cat((skips[-i - 1], x), 1), 1, False, ins[i] != outs[i], training, stats_out)
    pred = F.conv2d(x, params["layer_up_last.weight"], params["layer_up_last.bias"])
    seg = F.conv2d(x, params["output_layer.weight"], params["output_layer.bias"])
    sd = cfg.get("semantic_dim", 0)
    out = {"prediction": pred, "latents": latents, "segments": seg[:, sd:]}
    if sd > 0:
        out["semantics"] = seg[:, :sd]
    return out


# --------------------------------------------------------------------------------------------
# upfirdn2d reference (a'2)
# --------------------------------------------------------------------------------------------
def upfirdn2d_ref(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """lib/components/ops/upfirdn2d.py:165-211 (`_upfirdn2d_ref`): zero-insert up, pad/crop, FIR, decimate."""
    upx, upy = (up, up) if isinstance(up, int) else up
    downx, downy = (down, down) if isinstance(down, int) else down
    if isinstance(padding, int):
        padding = [padding] * 4
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    px0, px1, py0, py1 = padding
    B, C, H, W = x.shape
    if f is None:
        f = torch.ones(1, 1, dtype=torch.float32)
    x = x.reshape(B, C, H, 1, W, 1)
    x = F.pad(x, [0, upx - 1, 0, 0, 0, upy - 1])
    x = x.reshape(B, C, H * upy, W * upx)
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    f = f * (gain ** (f.ndim / 2))
    f = f.to(x.dtype)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f[None, None].repeat([C, 1] + [1] * f.ndim)
    if f.ndim == 4:
        x = F.conv2d(x, f, groups=C)
    else:
        x = F.conv2d(x, f.unsqueeze(2), groups=C)
        x = F.conv2d(x, f.unsqueeze(3), groups=C)
    return x[:, :, ::downy, ::downx]


# --------------------------------------------------------------------------------------------
# deterministic parameter initialisation (distributions follow the reference initialisers)
# --------------------------------------------------------------------------------------------
def init_generator_params(cfg, seed=0, sigma_gain=1.0, sigma_bias=0.0):
    """Seeded parameter dict with the reference's state_dict names/shapes.

    Distributions follow the reference initialisers (pigan_layers.py:20-52, util.py:7-10,
    map3d_layers.py:210,247,340, mapping_networks.py:28-30,105) but the draw order is this file's
    own, so values are NOT those of a seeded reference module: fixtures load THIS dict into the
    reference with load_state_dict.  `sigma_gain/bias` rescale the sigma head so that densities
    are non-trivial (SURVEY.md §8c pitfall 6)."""
    g = torch.Generator().manual_seed(seed)
    H, Fd, L = cfg["hidden_dim"], cfg["feature_dim"], cfg["latent_dim"]
    G, nb = cfg["geo_feature_dim"], cfg["neural_field_blocks"]
    p = {}
    U = lambda shape, a: (torch.rand(shape, generator=g) * 2 - 1) * a
    N = lambda shape, s=1.0: torch.randn(shape, generator=g) * s

    def lin(name, out_f, in_f, wbound, bias_bound=None):
        p[name + ".weight"] = U((out_f, in_f), wbound)
        p[name + ".bias"] = U((out_f,), bias_bound if bias_bound is not None else 1 / math.sqrt(in_f))

    lin("neural_field.first_layer_coord.layer", H, 3, 1 / 3)
    lin("neural_field.first_layer_mod.layer", H, G, 1 / G)
    lin("neural_field.network.0.layer", H, 2 * H, math.sqrt(6 / (2 * H)) / 25)
    for i in range(1, nb):
        lin(f"neural_field.network.{i}.layer", H, H, math.sqrt(6 / H) / 25)
    lin("neural_field.sigma_layer", 1, H, math.sqrt(6 / H) / 25)
    p["neural_field.sigma_layer.weight"] *= sigma_gain
    p["neural_field.sigma_layer.bias"] = p["neural_field.sigma_layer.bias"] * sigma_gain + sigma_bias
    lin("neural_field.color_layer_sine.layer", H, H + 3, math.sqrt(6 / (H + 3)) / 25)
    lin("neural_field.color_layer_linear", 3, H, math.sqrt(6 / H) / 25)
    lin("neural_field.feature_layer_linear", Fd, H, math.sqrt(6 / H) / 25)

    p["synthesis_input.network.0.weight"] = U((Fd, 2, 1, 1), math.sqrt(9 / 2))
    p["synthesis_input.network.0.bias"] = U((Fd,), 1 / math.sqrt(2))
    sin_in = 1 if "segments" in cfg["condition_modal_gen"] else 3
    p["synthesis_style_input.from_coords.0.weight"] = U((L, sin_in, 1, 1), math.sqrt(9 / sin_in))
    p["synthesis_style_input.from_coords.0.bias"] = U((L,), 1 / math.sqrt(sin_in))
    kstd = lambda fan_in: math.sqrt(2 / (1 + LRELU ** 2)) / math.sqrt(fan_in)
    p["synthesis_style_input.network.0.weight"] = N((Fd, 2 * L, 1, 1), kstd(2 * L))
    p["synthesis_style_input.network.0.bias"] = U((Fd,), 1 / math.sqrt(2 * L))
    p["synthesis_style_input.network.2.weight"] = U((Fd, Fd, 1, 1), 1 / math.sqrt(Fd))
    p["synthesis_style_input.network.2.bias"] = U((Fd,), 1 / math.sqrt(Fd))

    for k in range(cfg["synthesis_blocks"]):
        blk = f"synthesis_network.network.m3d_{k}."
        for c in ("conv_0.", "conv_1."):
            p[blk + c + "bias"] = U((H,), 1 / math.sqrt(H))
            p[blk + c + "weight_orig"] = N((H, H, 1, 1), 1 / math.sqrt(H))
            p[blk + c + "weight_u"] = F.normalize(N((H,)), dim=0)
            p[blk + c + "weight_v"] = F.normalize(N((H,)), dim=0)
        for s in ("spade_0.", "spade_1."):
            p[blk + s + "first_norm.weight"] = 1 + N((H,), 0.1)
            p[blk + s + "first_norm.bias"] = N((H,), 0.1)
            p[blk + s + "first_norm.running_mean"] = torch.zeros(H)
            p[blk + s + "first_norm.running_var"] = torch.ones(H)
            p[blk + s + "first_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
            p[blk + s + "mlp_shared.0.weight"] = U((128, Fd, 1, 1), 1 / math.sqrt(Fd))
            p[blk + s + "mlp_shared.0.bias"] = U((128,), 1 / math.sqrt(Fd))
            for m in ("mlp_gamma.", "mlp_beta."):
                p[blk + s + m + "weight"] = U((H, 128, 1, 1), 1 / math.sqrt(128))
                p[blk + s + m + "bias"] = U((H,), 1 / math.sqrt(128))
    for k in range(cfg["synthesis_blocks"]):
        p[f"synthesis_network.to_rgbs.m3d_{k}.linear.weight"] = U((3, H, 1, 1), 0.25 / math.sqrt(H))
        p[f"synthesis_network.to_rgbs.m3d_{k}.linear.bias"] = U((3,), 1 / math.sqrt(H))

    for i in (0, 2, 4):
        p[f"neural_field_mapping_network.network.{i}.weight"] = N((H, L if i == 0 else H), kstd(L if i == 0 else H))
        p[f"neural_field_mapping_network.network.{i}.bias"] = U((H,), 1 / math.sqrt(H))
    p["neural_field_mapping_network.network.6.weight"] = N((2 * nb * H, H), 0.25 * kstd(H))
    p["neural_field_mapping_network.network.6.bias"] = U((2 * nb * H,), 1 / math.sqrt(H))
    for i in range(7):
        p[f"synthesis_mapping_network.trunk{i}.weight"] = N((Fd, L if i == 0 else Fd), 100.0)
        p[f"synthesis_mapping_network.trunk{i}.bias"] = torch.zeros(Fd)
    p["synthesis_mapping_network.implicit0.weight"] = N((1, Fd), 100.0)
    p["synthesis_mapping_network.implicit0.bias"] = torch.zeros(1)
    p["synthesis_mapping_network.superres0.weight"] = N((Fd, Fd), 100.0)
    p["synthesis_mapping_network.superres0.bias"] = torch.zeros(Fd)
    p["latent_pool.latents"] = torch.zeros(cfg["dataset_length"], L)
    return p


def init_discriminator_params(cfg, seed=0):
    """Seeded D parameters with the reference's names/shapes (kaiming-normal a=0.2 fan_in, :74-79)."""
    g = torch.Generator().manual_seed(seed)
    N = lambda shape, s=1.0: torch.randn(shape, generator=g) * s
    U = lambda shape, a: (torch.rand(shape, generator=g) * 2 - 1) * a
    kstd = lambda fan_in: math.sqrt(2 / (1 + LRELU ** 2)) / math.sqrt(fan_in)
    nb = min(cfg.get("discriminator_blocks", 6), int(math.log2(max(cfg["gen_height"], cfg["gen_width"]))) - 1)
    ch = [3, 128, 128, 256, 256, 512, 512, 512, 512]
    p = {}

    def sn(name, co, ci, k):
        fan = ci * k * k
        p[name + ".bias"] = U((co,), 1 / math.sqrt(fan))
        p[name + ".weight_orig"] = N((co, ci, k, k), kstd(fan))
        p[name + ".weight_u"] = F.normalize(N((co,)), dim=0)
        p[name + ".weight_v"] = F.normalize(N((fan,)), dim=0)

    def block(name, fin, fout, first, up):
        c1 = ".conv1" if first else (".conv1.2" if up else ".conv1.1")
        sn(name + c1, fout, fin, 3)
        sn(name + ".conv2.1", fout, fout, 3)
        if fin != fout:
            sn(name + ".conv_s", fout, fin, 1)

    outs = [ch[nb - 1]] + [ch[nb - i - 1] for i in range(1, nb - 1)] + [64]
    ins = [ch[nb]] + [2 * ch[nb - i] for i in range(1, nb - 1)] + [2 * ch[1]]
    for i in range(nb):
        block(f"body_up.{i}", ins[i], outs[i], False, True)
    for i in range(nb):
        block(f"body_down.{i}", ch[i], ch[i + 1], i == 0, False)
    od = cfg.get("semantic_dim", 0) + cfg.get("label_dim", 0)
    p["layer_up_last.weight"] = N((1, 64, 1, 1), kstd(64))
    p["layer_up_last.bias"] = U((1,), 1 / 8)
    p["output_layer.weight"] = N((od, 64, 1, 1), 0.25 * kstd(64))
    p["output_layer.bias"] = U((od,), 1 / 8)
    ds = 2 ** nb
    kh, kw = cfg["gen_height"] // ds, cfg["gen_width"] // ds
    p["latent_layer.weight"] = N((cfg["latent_dim"], ch[nb], kh, kw), kstd(ch[nb] * kh * kw))
    p["latent_layer.bias"] = U((cfg["latent_dim"],), 1 / math.sqrt(ch[nb] * kh * kw))
    return p
