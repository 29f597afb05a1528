"""CPU oracle for the 3D U-Net hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (pytorch3dunet_b200/) never does; it fails loudly when
its CUDA library is missing.

This is a *functional restatement* (plain torch CPU ops over a flat state_dict) of the
reference's model path.  All file:line citations are relative to /root/reference/:

  pytorch3dunet/unet3d/model.py          AbstractUNet.__init__ :38-101, _forward_logits :123-149,
                                         UNet3D :152, ResidualUNet3D :193, ResidualUNetSE3D :237
  pytorch3dunet/unet3d/buildingblocks.py create_conv :10-96, SingleConv :99, DoubleConv :138-227,
                                         ResNetBlock :230-288, ResNetBlockSE :291-307,
                                         Encoder :310-384, Decoder :387-493,
                                         create_encoders :496-544, create_decoders :547-574,
                                         InterpolateUpsampling :598-614, TransposeConvUpsampling :617-664
  pytorch3dunet/unet3d/se.py             ChannelSELayer3D :18-51, SpatialSELayer3D :54-93,
                                         ChannelSpatialSELayer3D :96-114
  pytorch3dunet/unet3d/utils.py          number_of_features_per_level :110-112
  pytorch3dunet/unet3d/losses.py         compute_per_channel_dice :11-37, DiceLoss :130-145,
                                         BCEDiceLoss :187-201, flatten :253-271

The arithmetic itself lives in PyTorch/ATen (un-pinned third-party dependency of the reference,
pyproject.toml:17); installed here: torch 2.11.0.  Parity pinning: tests/test_oracle.py checks this
restatement against golden vectors produced by the *reference's own classes* run in the build
container (oracle/make_golden.py, fixtures in tests/golden/*.npz).  The reference's own tests hold no
numeric vectors for this path, so those fixtures are the pin.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GN_EPS = 1e-5  # torch.nn.GroupNorm default, buildingblocks.py:75 passes none


# --------------------------------------------------------------------------------------
# architecture plan (pure python, no tensors)
# --------------------------------------------------------------------------------------
def features_per_level(f_maps, num_levels):
    """utils.py:110-112 : f_maps * 2**k for k in range(num_levels)."""
    if isinstance(f_maps, int):
        return [f_maps * 2 ** k for k in range(num_levels)]
    return list(f_maps)


MODEL_DEFAULTS = {
    # model.py:159-174 / :203-218 / :247-262
    "UNet3D": dict(basic="double", num_levels=4),
    "ResidualUNet3D": dict(basic="res", num_levels=5),
    "ResidualUNetSE3D": dict(basic="res_se", num_levels=5),
}


def normalize_config(cfg):
    """Fill in the constructor defaults of the named 3-D model class."""
    name = cfg["name"]
    d = MODEL_DEFAULTS[name]
    out = dict(
        name=name,
        in_channels=cfg["in_channels"],
        out_channels=cfg["out_channels"],
        final_sigmoid=cfg.get("final_sigmoid", True),
        f_maps=cfg.get("f_maps", 64),
        layer_order=cfg.get("layer_order", "gcr"),
        num_groups=cfg.get("num_groups", 8),
        num_levels=cfg.get("num_levels", d["num_levels"]),
        is_segmentation=cfg.get("is_segmentation", True),
        conv_padding=cfg.get("conv_padding", 1),
        conv_upscale=cfg.get("conv_upscale", 2),
        upsample=cfg.get("upsample", "default"),
        basic=d["basic"],
    )
    out["f_maps"] = features_per_level(out["f_maps"], out["num_levels"])
    return out


# --------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------
def _groups(num_channels, num_groups):
    # buildingblocks.py:69-70 : a single group when there are fewer channels than groups
    return 1 if num_channels < num_groups else num_groups


def single_conv(x, sd, prefix, order, num_groups, padding=1, masks=None):
    """create_conv / SingleConv, buildingblocks.py:10-135.  `order` is walked left to right.

    masks (tests only): {prefix: bool NCDHW tensor}.  When given, ReLU is evaluated with THAT activation pattern
    (x * mask) instead of x > 0 -- lets a gradient parity test run the reference arithmetic at the activation
    pattern of the implementation under test, so that bf16-induced flips of near-zero ReLU inputs (a
    discontinuity, not an arithmetic error) do not dominate the comparison."""
    for ch in order:
        if ch == "c":
            x = F.conv3d(x, sd[prefix + "conv.weight"], sd.get(prefix + "conv.bias"), padding=padding)
        elif ch == "g":
            w = sd[prefix + "groupnorm.weight"]
            x = F.group_norm(x, _groups(w.numel(), num_groups), w, sd[prefix + "groupnorm.bias"], GN_EPS)
        elif ch == "r":
            if masks is not None and "__record__" in masks:  # tests: capture this run's own activation pattern (flip-rate checks)
                masks["__record__"][prefix] = x > 0
            x = x * masks[prefix].to(x.dtype) if (masks is not None and prefix in masks) else F.relu(x)
        elif ch == "l":  # nn.LeakyReLU() default slope, buildingblocks.py:49 (the kink is pinned like ReLU's when masks are given)
            x = torch.where(masks[prefix], x, 0.01 * x) if (masks is not None and prefix in masks) else F.leaky_relu(x, 0.01)
        elif ch == "e":
            x = F.elu(x)
        elif ch == "b":
            x = F.batch_norm(x, sd[prefix + "batchnorm.running_mean"], sd[prefix + "batchnorm.running_var"],
                             sd[prefix + "batchnorm.weight"], sd[prefix + "batchnorm.bias"], False, 0.1, 1e-5)
        elif ch in "dD":
            pass  # dropout is identity in the deterministic oracle (eval semantics)
        else:
            raise ValueError(ch)
    return x


def double_conv(x, sd, prefix, order, num_groups, padding=1, masks=None):
    """DoubleConv.forward = SingleConv1 then SingleConv2 (nn.Sequential), buildingblocks.py:200-227."""
    x = single_conv(x, sd, prefix + "SingleConv1.", order, num_groups, padding, masks)
    return single_conv(x, sd, prefix + "SingleConv2.", order, num_groups, padding, masks)


def channel_se(x, sd, prefix):
    """se.py:40-51."""
    n, c = x.shape[:2]
    s = x.mean(dim=(2, 3, 4))
    h = F.relu(F.linear(s, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]))
    g = torch.sigmoid(F.linear(h, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"]))
    return x * g.view(n, c, 1, 1, 1)


def spatial_se(x, sd, prefix):
    """se.py:69-93 (the `weights` few-shot branch is never used by the model)."""
    g = torch.sigmoid(F.conv3d(x, sd[prefix + "conv.weight"], sd[prefix + "conv.bias"]))
    return x * g


def res_block(x, sd, prefix, order, num_groups, se=False, masks=None):
    """ResNetBlock.forward buildingblocks.py:277-288 (+ ResNetBlockSE :304-307).  `masks`: see single_conv; the block's
    final ReLU (after the residual add) is keyed by the conv3 prefix."""
    if (prefix + "conv1.weight") in sd:
        residual = F.conv3d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"])
    else:
        residual = x
    out = single_conv(residual, sd, prefix + "conv2.", order, num_groups, masks=masks)
    n_order = order.replace("r", "").replace("e", "").replace("l", "")
    out = single_conv(out, sd, prefix + "conv3.", n_order, num_groups, masks=masks)
    out = out + residual
    if "l" in order:  # buildingblocks.py:271 : slope 0.1 here, not 0.01
        key3 = prefix + "conv3."
        out = torch.where(masks[key3], out, 0.1 * out) if (masks is not None and key3 in masks) else F.leaky_relu(out, 0.1)
    elif "e" in order:
        out = F.elu(out)
    elif masks is not None and (prefix + "conv3.") in masks:
        out = out * masks[prefix + "conv3."].to(out.dtype)
    else:
        if masks is not None and "__record__" in masks:
            masks["__record__"][prefix + "conv3."] = out > 0
        out = F.relu(out)
    if se:
        cse = channel_se(out, sd, prefix + "se_module.cSE.")
        sse = spatial_se(out, sd, prefix + "se_module.sSE.")
        key = prefix + "se_module.#select"
        if masks is not None and key in masks:
            # tests only: which branch of the element-wise max wins is a discontinuity for the gate-parameter gradients
            # (g ~ q ~ 0.5 at initialisation, so bf16 noise flips many selections); evaluate at the given selection
            out = torch.where(masks[key], cse, sse)
        else:
            out = torch.max(cse, sse)  # se.py:113
    return out


def basic_module(x, sd, prefix, cfg, masks=None):
    if cfg["basic"] == "double":
        return double_conv(x, sd, prefix, cfg["layer_order"], cfg["num_groups"], cfg["conv_padding"], masks)
    return res_block(x, sd, prefix, cfg["layer_order"], cfg["num_groups"], se=cfg["basic"] == "res_se", masks=masks)


def decoder_mode(cfg):
    """Decoder.__init__ buildingblocks.py:431-468 -> (upsample kind, concat?)."""
    up = cfg["upsample"]
    concat = True
    if up == "default":
        if cfg["basic"] == "double":
            up, concat = "nearest", True
        else:
            up, concat = "deconv", False
    return up, concat


def max_pool_at(x, idx):
    """MaxPool3d(2) whose selected element per window is given (tests only, same purpose as `masks`): the argmax of a
    window with two nearly equal values is a discontinuity that bf16 rounding flips."""
    return x.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)


def forward(sd, cfg, x, masks=None, pool_idx=None):
    """AbstractUNet._forward_logits, model.py:123-149.  Returns (output, logits).  `masks`: see single_conv;
    `pool_idx`: list of max-pool argmax index tensors (one per pooling level), see max_pool_at."""
    cfg = normalize_config(cfg)
    nlev = len(cfg["f_maps"])
    feats = []
    for i in range(nlev):
        if i > 0:
            # Encoder.forward :380-384, MaxPool3d(kernel_size=2) :356
            x = max_pool_at(x, pool_idx[i - 1]) if pool_idx is not None else F.max_pool3d(x, 2)
        x = basic_module(x, sd, f"encoders.{i}.basic_module.", cfg, masks)
        feats.insert(0, x)
    feats = feats[1:]
    up, concat = decoder_mode(cfg)
    for i, enc in enumerate(feats):
        size = enc.shape[2:]
        if up == "deconv":
            # TransposeConvUpsampling.Upsample.forward :649-651 : deconv(k3,s2,p1) then nearest resize
            x = F.conv_transpose3d(x, sd[f"decoders.{i}.upsampling.upsample.conv_transposed.weight"],
                                   None, stride=2, padding=1)
            x = F.interpolate(x, size=size)
        elif up is None or up == "none":
            pass
        else:
            x = F.interpolate(x, size=size, mode=up)  # InterpolateUpsampling._interpolate :613-614
        x = torch.cat((enc, x), dim=1) if concat else enc + x  # Decoder._joining :488-493
        x = basic_module(x, sd, f"decoders.{i}.basic_module.", cfg, masks)
    logits = F.conv3d(x, sd["final_conv.weight"], sd["final_conv.bias"])  # model.py:89,141
    if cfg["is_segmentation"]:
        out = torch.sigmoid(logits) if cfg["final_sigmoid"] else torch.softmax(logits, dim=1)
        return out, logits
    return logits, logits


# --------------------------------------------------------------------------------------
# parameter shapes (state_dict contract)
# --------------------------------------------------------------------------------------
def _single_conv_shapes(shapes, prefix, cin, cout, order, k=3):
    for i, ch in enumerate(order):
        if ch == "c":
            shapes[prefix + "conv.weight"] = (cout, cin, k, k, k)
            if not ("g" in order or "b" in order):
                shapes[prefix + "conv.bias"] = (cout,)
        elif ch == "g":
            c = cin if i < order.index("c") else cout
            shapes[prefix + "groupnorm.weight"] = (c,)
            shapes[prefix + "groupnorm.bias"] = (c,)


def _basic_shapes(shapes, prefix, cin, cout, cfg, encoder):
    order = cfg["layer_order"]
    if cfg["basic"] == "double":
        if encoder:  # buildingblocks.py:177-186
            c1 = cout if cfg["conv_upscale"] == 1 else cout // 2
            if c1 < cin:
                c1 = cin
        else:
            c1 = cout
        _single_conv_shapes(shapes, prefix + "SingleConv1.", cin, c1, order)
        _single_conv_shapes(shapes, prefix + "SingleConv2.", c1, cout, order)
    else:
        if cin != cout:
            shapes[prefix + "conv1.weight"] = (cout, cin, 1, 1, 1)
            shapes[prefix + "conv1.bias"] = (cout,)
        _single_conv_shapes(shapes, prefix + "conv2.", cout, cout, order)
        n_order = order.replace("r", "").replace("e", "").replace("l", "")
        _single_conv_shapes(shapes, prefix + "conv3.", cout, cout, n_order)
        if cfg["basic"] == "res_se":
            p = prefix + "se_module."
            shapes[p + "cSE.fc1.weight"] = (cout, cout)
            shapes[p + "cSE.fc1.bias"] = (cout,)
            shapes[p + "cSE.fc2.weight"] = (cout, cout)
            shapes[p + "cSE.fc2.bias"] = (cout,)
            shapes[p + "sSE.conv.weight"] = (1, cout, 1, 1, 1)
            shapes[p + "sSE.conv.bias"] = (1,)


def param_shapes(cfg):
    cfg = normalize_config(cfg)
    f = cfg["f_maps"]
    shapes = {}
    for i, fo in enumerate(f):
        cin = cfg["in_channels"] if i == 0 else f[i - 1]
        _basic_shapes(shapes, f"encoders.{i}.basic_module.", cin, fo, cfg, True)
    up, concat = decoder_mode(cfg)
    rf = f[::-1]
    for i in range(len(rf) - 1):
        # create_decoders :553-559
        cin = rf[i] + rf[i + 1] if (cfg["basic"] == "double" and up != "deconv") else rf[i]
        cout = rf[i + 1]
        if up == "deconv":
            shapes[f"decoders.{i}.upsampling.upsample.conv_transposed.weight"] = (cin, cout, 3, 3, 3)
        if cfg["basic"] != "double" and cfg["upsample"] == "default":
            cin = cout  # adapt_channels, Decoder.__init__ :466-468
        _basic_shapes(shapes, f"decoders.{i}.basic_module.", cin, cout, cfg, False)
    shapes["final_conv.weight"] = (cfg["out_channels"], f[0], 1, 1, 1)
    shapes["final_conv.bias"] = (cfg["out_channels"],)
    return shapes


def random_state_dict(cfg, seed=0, dtype=torch.float32):
    """Seeded synthetic weights with the reference's shapes (not the reference's init distribution)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k.endswith("groupnorm.weight"):
            t = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            t = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            t = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
        sd[k] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# losses (losses.py) -- used to close the fwd+bwd loop exactly like trainer.py:364-365
# --------------------------------------------------------------------------------------
def _flatten(t):
    """losses.py:253-271 : (N,C,D,H,W) -> (C, N*D*H*W)."""
    c = t.size(1)
    return t.transpose(0, 1).reshape(c, -1)


def dice_loss(logits, target, eps=1e-6):
    """DiceLoss with sigmoid normalisation, losses.py:130-145 + :11-37."""
    p = _flatten(torch.sigmoid(logits))
    t = _flatten(target).float()
    inter = (p * t).sum(-1)
    den = (p * p).sum(-1) + (t * t).sum(-1)
    return 1.0 - (2 * (inter / den.clamp(min=eps))).mean()


def bce_dice_loss(logits, target, alpha=1.0):
    """BCEDiceLoss, losses.py:187-201."""
    return F.binary_cross_entropy_with_logits(logits, target) + alpha * dice_loss(logits, target)
