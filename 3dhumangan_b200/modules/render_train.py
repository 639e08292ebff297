"""Training-mode forward + backward of the FiLM-SIREN renderer (COORDCONCATSIREN.forward modulated.py:41-75 +
vr.ray_integration volume_rendering.py:12-56) on the sm_100a kernels.

The inference path is ONE fused kernel (csrc/render.cu) that keeps every activation on chip.  Training needs the
pre-activations back, so here the MLP runs layer by layer over tile-blocked points [B, T, 256, 128] (the layout of the
synthesis network; a "pixel" is a sample point p = ray*S + s) and keeps the 8 linear outputs:

    lin_a = Wc x + bc, lin_b = Wg g + bg                         hg_conv1x1_blocked        (K = 3 / 31, zero padded)
    out_0 = W0 [sin(30 lin_a); sin(30 lin_b)] + b0               hg_act_conv1x1_blocked    (sine operand, K = 512)
    out_i = W_i sin(f_{i-1} out_{i-1} + phi_{i-1}) + b_i         i = 1..3
    lin_c = Wcol[:,3:] sin(f_3 out_3 + phi_3) + (bcol + Wcol[:,:3] dir)
    feat  = Wf sin(f_3 lin_c + phi_3) + bf
    sigma, rgb_pre                                               hg_render_heads
    ray_out = composite(...)                                     hg_render_composite

Backward: hg_render_composite_bwd, hg_render_heads_bwd, then per layer the blocked data-gradient kernel with the
cosine mask (the sigma / rgb heads enter as rank-1 / rank-3 terms of its epilogue, the FiLM frequency of the consumer
as a per-(sample, channel) scale of its operand) and the blocked weight-gradient kernel with the sine operand.
d freq / d phase come from the per-(b,c) sums S1 = sum dpre, S2 = sum dpre*x of the data-gradient kernels through
torch autograd on the [B,256] tables.  Geometry features carry no gradient (wrapped in no_grad in the reference,
map3d_generator.py:196-205).
"""
from __future__ import annotations

import torch

from .. import abi

H = 256


def blocked_points(t, C):
    """[B,N,c] point-major -> tile-blocked [B,T,C,128] (channels zero padded to C)."""
    B, N, c = t.shape
    assert N % 128 == 0
    out = torch.zeros(B, N // 128, C, 128, dtype=torch.float32, device=t.device)
    out[:, :, :c] = t.reshape(B, N // 128, 128, c).permute(0, 1, 3, 2)
    return out


def mlp_forward_train(P, freq, phase, rec, z_vals, noise, cfg, *, geo_dim=31, locked_dir=(0.0, 0.0, -1.0), passes=3,
                      prefix="neural_field."):
    """rec [B,N,>=3+geo_dim] (scaled coordinates, geometry features), z_vals [B,N], noise [B,N] or None.
    Returns (ray_out [B,R,260] without autograd history, tape)."""
    abi.require_device()
    g = lambda n: P[prefix + n]
    dev = rec.device
    B, N = rec.shape[0], rec.shape[1]
    S = cfg["num_steps"]
    R = N // S
    if cfg.get("last_back", False):
        raise RuntimeError("hg3d: last_back=True is an inference-only setting (eval_last_back); the training renderer does not build it")
    if g("network.0.layer.weight").shape[0] != H:
        raise RuntimeError("hg3d: the sm_100a render kernels are built for hidden_dim == 256")
    f32 = dict(dtype=torch.float32, device=dev)
    kw = dict(B=B, Hg=1, Wg=N, passes=passes)
    pack = lambda w: abi.pack_weight(w.detach().float().contiguous(), Nb=256)[0]

    # FiLM tables with autograd history: f = 15*freq + 30 (modulated.py:43); the colour layer re-uses the last slice
    fq = freq.detach().float().requires_grad_(True)
    ph = phase.detach().float().requires_grad_(True)
    f = fq * 15 + 30
    mods = [torch.stack([f[:, i * H:(i + 1) * H], ph[:, i * H:(i + 1) * H]], dim=1) for i in range(4)]      # [B,2,256] each
    mod30 = torch.stack([torch.full((B, H), 30.0, **f32), torch.zeros(B, H, **f32)], dim=1).contiguous()
    mods_d = [m.detach().contiguous() for m in mods]

    rec_b = blocked_points(rec[..., :3 + geo_dim], 128)
    Wa = torch.zeros(H, 128, **f32)
    Wb = torch.zeros(H, 128, **f32)
    Wa[:, :3] = g("first_layer_coord.layer.weight").detach()
    Wb[:, 3:3 + geo_dim] = g("first_layer_mod.layer.weight").detach()
    new = lambda: torch.empty(B, N // 128, H, 128, **f32)
    lin_a = abi.conv1x1_blocked(rec_b, 128, pack(Wa), g("first_layer_coord.layer.bias").detach(), new(), **kw)
    lin_b = abi.conv1x1_blocked(rec_b, 128, pack(Wb), g("first_layer_mod.layer.bias").detach(), new(), **kw)
    outs = [abi.act_conv1x1_blocked(lin_a, mod30, pack(g("network.0.layer.weight")), g("network.0.layer.bias").detach(), new(),
                                    x2=lin_b, **kw)]
    for i in range(1, 4):
        outs.append(abi.act_conv1x1_blocked(outs[-1], mods_d[i - 1], pack(g(f"network.{i}.layer.weight")),
                                            g(f"network.{i}.layer.bias").detach(), new(), **kw))
    wcol = g("color_layer_sine.layer.weight")
    dvec = torch.tensor(locked_dir, **f32)
    bcol = g("color_layer_sine.layer.bias") + wcol[:, :3] @ dvec            # autograd: bias and the direction columns
    lin_c = abi.act_conv1x1_blocked(outs[3], mods_d[3], pack(wcol[:, 3:]), bcol.detach().contiguous(), new(), **kw)
    feat = abi.act_conv1x1_blocked(lin_c, mods_d[3], pack(g("feature_layer_linear.weight")),
                                   g("feature_layer_linear.bias").detach(), new(), **kw)
    w_sigma = g("sigma_layer.weight").detach().reshape(-1).float().contiguous()
    w_rgb = g("color_layer_linear.weight").detach().float().contiguous()
    heads_b = torch.cat([g("sigma_layer.bias").detach().reshape(1), g("color_layer_linear.bias").detach().reshape(3)]).float().contiguous()
    sig, rgbp = abi.render_heads(outs[3], lin_c, mods_d[3], w_sigma, w_rgb, heads_b, B=B, N=N)
    noise = None if noise is None else noise.reshape(B, N).float().contiguous()
    comp = dict(B=B, R=R, S=S, noise_std=cfg["nerf_noise"], white_back=cfg.get("white_back", False),
                softplus=cfg["clamp_mode"] == "softplus")
    ray_out, w = abi.render_composite(sig, z_vals, noise, rgbp, feat, **comp)
    tape = dict(P=P, prefix=prefix, geo_dim=geo_dim, fq=fq, ph=ph, mods=mods, mods_d=mods_d, mod30=mod30, rec_b=rec_b,
                lin_a=lin_a, lin_b=lin_b, outs=outs, lin_c=lin_c, feat=feat, sig=sig, rgbp=rgbp, z=z_vals, noise=noise, comp=comp,
                kw=kw, bcol=bcol, w_sigma=w_sigma, w_rgb=w_rgb, B=B, N=N, weights=w)
    return ray_out, tape


def mlp_backward(tape, dray, grads=None):
    """dray [B,R,260] (gradient w.r.t. feat | rgb | depth; the depth column is ignored, as no loss uses it).
    Adds the gradients of the neural-field parameters to `grads` (name -> tensor; to `.grad` when grads is None) and
    returns (d freq, d phase) [B,4*256] each."""
    from .synthesis_train import grad_accumulator
    P, prefix = tape["P"], tape["prefix"]
    g = lambda n: P[prefix + n]
    acc_ = grad_accumulator(P, grads)
    acc = lambda n, grad: acc_(prefix + n, grad)
    B, N, kw = tape["B"], tape["N"], tape["kw"]
    dev = dray.device
    f32 = dict(dtype=torch.float32, device=dev)
    T = N // 128
    full = T * H * 128
    new = lambda: torch.empty(B, T, H, 128, **f32)
    packT = lambda w: abi.pack_weight(w.detach().float().t().contiguous(), Nb=256)[0]
    mods_d, mod30, outs = tape["mods_d"], tape["mod30"], tape["outs"]

    dfeat, drgbp, dsig = abi.render_composite_bwd(tape["sig"], tape["z"], tape["noise"], tape["rgbp"], tape["feat"],
                                                  dray.float().contiguous(), **tape["comp"])
    hb = abi.render_heads_bwd(outs[3], tape["lin_c"], mods_d[3], dsig, drgbp, B=B, N=N)
    acc("sigma_layer.weight", hb[:H].float())
    acc("color_layer_linear.weight", hb[H:4 * H].float().reshape(3, H))
    acc("sigma_layer.bias", hb[4 * H:4 * H + 1].float())
    acc("color_layer_linear.bias", hb[4 * H + 1:].float())

    dmods = [torch.zeros(B, 2, H, **f32) for _ in range(4)]
    sums = lambda: torch.zeros(B, 2, H, dtype=torch.float64, device=dev)

    def add_film(i, s):          # d g1 = sum dpre*x, d g0 = sum dpre
        dmods[i] += torch.stack([s[:, 1], s[:, 0]], dim=1).float()

    rk3 = tape["w_rgb"]                                                        # [3,256]
    rk1 = torch.zeros(3, H, **f32)
    rk1[0] = tape["w_sigma"]
    # ---- feature layer: feat = Wf sin(f3 lin_c + phi3) + bf;  the rgb head feeds back through the same activation
    s = sums()
    dpre_c = abi.conv1x1_blocked_bwd(dfeat, tape["lin_c"], packT(g("feature_layer_linear.weight")), new(), s, mod=mods_d[3], act=1,
                                     rk_w=rk3, rk_v=drgbp, **kw)
    add_film(3, s)
    dw, db = abi.act_wgrad_blocked(dfeat, tape["lin_c"], full, mods_d[3], act=1, **kw)
    acc("feature_layer_linear.weight", dw)
    acc("feature_layer_linear.bias", db)
    del dfeat
    # ---- colour layer: lin_c = Wcol' sin(f3 out3 + phi3) + bcol';  the sigma head feeds back through h4
    f3 = mods_d[3][:, 0].contiguous()
    s = sums()
    wcol = g("color_layer_sine.layer.weight")
    dpre = abi.conv1x1_blocked_bwd(dpre_c, outs[3], packT(wcol[:, 3:]), new(), s, mod=mods_d[3], act=1, ascale=f3,
                                   rk_w=rk1, rk_v=dsig.reshape(B, 1, N), **kw)
    add_film(3, s)
    dw, db = abi.act_wgrad_blocked(dpre_c, outs[3], full, mods_d[3], act=1, pscale=f3, **kw)
    gw = torch.zeros_like(wcol)
    gw[:, 3:] = dw
    acc("color_layer_sine.layer.weight", gw)
    small = [(tape["bcol"], db)]                                              # bias + direction columns via autograd
    del dpre_c
    # ---- network.3 .. network.1: out_i = W_i sin(f_{i-1} out_{i-1} + phi_{i-1}) + b_i
    for i in (3, 2, 1):
        fi = mods_d[i][:, 0].contiguous()
        s = sums()
        wi = g(f"network.{i}.layer.weight")
        nxt = abi.conv1x1_blocked_bwd(dpre, outs[i - 1], packT(wi), new(), s, mod=mods_d[i - 1], act=1, ascale=fi, **kw)
        add_film(i - 1, s)
        dw, db = abi.act_wgrad_blocked(dpre, outs[i - 1], full, mods_d[i - 1], act=1, pscale=fi, **kw)
        acc(f"network.{i}.layer.weight", dw)
        acc(f"network.{i}.layer.bias", db)
        dpre = nxt
    # ---- network.0 (K = 512: coordinate half, geometry half) and the two first layers
    f0 = mods_d[0][:, 0].contiguous()
    w0 = g("network.0.layer.weight")
    gw0 = torch.empty(H, 2 * H, **f32)
    thirty = mod30[:, 0].contiguous()
    for half, lin, first, cols in ((0, tape["lin_a"], "first_layer_coord.layer.", slice(0, 3)),
                                   (1, tape["lin_b"], "first_layer_mod.layer.", slice(3, 3 + tape["geo_dim"]))):
        s = sums()
        dlin = abi.conv1x1_blocked_bwd(dpre, lin, packT(w0[:, half * H:(half + 1) * H]), new(), s, mod=mod30, act=1, ascale=f0, **kw)
        dw, db0 = abi.act_wgrad_blocked(dpre, lin, full, mod30, act=1, pscale=f0, **kw)
        gw0[:, half * H:(half + 1) * H] = dw
        # first layer: lin = W x + b with the sine's factor 30 folded into the incoming gradient
        dwf, dbf = abi.act_wgrad_blocked(dlin, tape["rec_b"], T * 128 * 128, None, act=2, pscale=thirty, Cx=128, **kw)
        acc(first + "weight", dwf[:, cols])
        acc(first + "bias", dbf)
        del dlin
    acc("network.0.layer.weight", gw0)
    acc("network.0.layer.bias", db0)
    # ---- FiLM tables, colour bias / direction columns: tiny autograd graphs
    outs_, grads_ = list(tape["mods"]), list(dmods)
    for t, gr in small:
        if t.requires_grad:
            outs_.append(t)
            grads_.append(gr)
    leaves = [n for n in ("color_layer_sine.layer.bias", "color_layer_sine.layer.weight") if g(n).requires_grad]
    res = torch.autograd.grad(outs_, [tape["fq"], tape["ph"]] + [g(n) for n in leaves], grads_, allow_unused=True)
    for n, r in zip(leaves, res[2:]):
        if r is not None:
            acc(n, r)
    return res[0], res[1]


@torch.no_grad()
def geo_records(cond, cfg, u):
    """Ray sampling + jitter + camera transform + exact nearest vertex + 31-d features (no gradient, as in the
    reference: map3d_generator.py:196-205) -> (rec [B,N,36], z_vals [B,N]); same launches as render_ops.render_forward."""
    dev = cond["vertices"].device
    B = cond["vertices"].shape[0]
    Rw, Rh, S = cfg["render_width"], cfg["render_height"], cfg["num_steps"]
    f32 = dict(dtype=torch.float32, device=dev)
    xs = torch.linspace(-Rw / Rh, Rw / Rh, Rw, **f32)
    ys = torch.linspace(-1, 1, Rh, **f32)
    zs = torch.linspace(cfg["ray_start"], cfg["ray_end"], S, **f32)
    vik = abi.vertex_ik(cond["fk_matrices"], cond["lbs_weights"])
    geo = abi.geo_features(cond["vertices"], cond["tpose_vertices"], cond["skeletons_xyz"], vik,
                           input_scaler=2.0 / cfg["side_length"], legacy_mode=cfg.get("legacy_mode", False),
                           xs=xs, ys=ys, zs=zs, focals=cond["intrinsics"][:, 0, 0], scales=cond["scales"],
                           cam2world=cond["cam2world_matrices"], jitter=u.reshape(B, Rw * Rh * S) if u is not None else None)
    return geo["rec"], geo["z_vals"]


CORE_PREFIXES = ("neural_field.", "synthesis_network.", "synthesis_input.")


def core_parameters(module):
    """(names, tensors): the renderer / synthesis parameters that enter `GeneratorCore` as autograd inputs."""
    names, tensors = [], []
    for n, p in module.named_parameters():
        if n.startswith(CORE_PREFIXES):
            names.append(n)
            tensors.append(p)
    return names, tensors


class GeneratorCore(torch.autograd.Function):
    """(freq, phase, fixed style, *renderer and synthesis parameters) -> (rgbs, rgbs_render, depth) on the sm_100a kernels.

    EVERY parameter the kernels read is an input of this node and its gradient is RETURNED by `backward`, so the
    reference trainer's machinery sees them like any other autograd node's: `DistributedDataParallel` reducer hooks fire
    (base_trainer.py:102-104, find_unused_parameters=True walks the graph to them), `torch.autograd.grad(loss, params)`
    works, `GradScaler.unscale_` finds fp32 `.grad`s.  Under `torch.autocast` the inputs are taken as fp32 (the kernels
    compute in fp32 / bf16x3 whatever the ambient autocast dtype; phase_trainer.py:355,396,462 run the models under
    fp16 autocast).  First-order only: the reference never differentiates the generator twice (R1 acts on real images)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, module, cond, cfg, u, noise, passes, names, freq, phase, styles, *tensors):
        from . import synthesis_train
        P = module._params()
        for n, t in zip(names, tensors):           # the tensors autograd tracks (identical objects unless a caller wraps them)
            P[n] = t
        B = freq.shape[0]
        Rh, Rw = cfg["render_height"], cfg["render_width"]
        if cfg.get("hierarchical_sample", False) or not cfg.get("lock_view_dependence", False):
            raise RuntimeError("hg3d: the training renderer is built for hierarchical_sample=False, lock_view_dependence=True")
        if cfg.get("neural_field_blocks", 4) != 4:
            raise RuntimeError("hg3d: the training renderer is built for neural_field_blocks == 4 (all shipped curricula)")
        rec, z_vals = geo_records(cond, cfg, u)
        with torch.enable_grad():
            ray, rtape = mlp_forward_train(P, freq, phase, rec, z_vals, noise, cfg, geo_dim=cfg["geo_feature_dim"],
                                           passes=passes)
            rgb, stape = synthesis_train.synthesis_forward_train(P, ray, styles.reshape(B, -1), cfg, passes=passes)
        ctx.tapes = (P, rtape, stape, cfg, passes, names)
        rgb_render = (ray[..., 256:259] * 2 - 1).reshape(B, Rh, Rw, 3).permute(0, 3, 1, 2).contiguous()
        depth = ray[..., 259:260].contiguous()
        ctx.mark_non_differentiable(depth)
        return rgb, rgb_render, depth

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_rgb, d_rgb_render, d_depth):
        from . import synthesis_train
        if ctx.tapes is None:
            raise RuntimeError("hg3d: the generator's activation tape was released by a previous backward pass "
                               "(one backward per forward; retain_graph=True keeps the graph, not the tape)")
        P, rtape, stape, cfg, passes, names = ctx.tapes
        B = stape.B
        Rh, Rw = cfg["render_height"], cfg["render_width"]
        grads = {}
        with torch.enable_grad():
            dfs, dfeat = synthesis_train.synthesis_backward(P, stape, d_rgb, passes=passes, grads=grads)
        dray = torch.zeros(B, Rh * Rw, 260, dtype=torch.float32, device=d_rgb.device)
        if dfeat is not None:
            dray[..., :256] = dfeat
        if d_rgb_render is not None:
            dray[..., 256:259] = 2.0 * d_rgb_render.permute(0, 2, 3, 1).reshape(B, Rh * Rw, 3)
        with torch.enable_grad():
            dfreq, dphase = mlp_backward(rtape, dray, grads=grads)
        ctx.tapes = None
        # Every returned gradient must own its storage: autograd's AccumulateGrad steals a returned tensor as `.grad` when
        # nobody else holds the TensorImpl, so two parameters whose gradients are views of one buffer (the nine ToRGB biases
        # all receive sum(d_rgb)) would end up with ALIASED `.grad`s -- and every in-place pass over the gradients
        # (GradScaler.unscale_, clip_grad_norm_) would then hit that buffer once per alias, concurrently in the foreach kernels.
        out, seen = [], set()
        for n in names:
            g = grads.get(n)
            if g is not None:
                g = g.detach()
                key = g.untyped_storage().data_ptr()
                if key in seen or g.untyped_storage().nbytes() != g.numel() * g.element_size():
                    g = g.clone()
                seen.add(g.untyped_storage().data_ptr())
            out.append(g)
        return (None, None, None, None, None, None, None, dfreq.detach(), dphase.detach(), dfs.detach().reshape(B, -1), *out)
