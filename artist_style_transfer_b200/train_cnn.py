"""B200-native mirror of the loss side of `/root/reference/train_cnn.py`.

Drop-in pieces (same names / signatures / return types as the reference):
  VGG16(just_content=False, vgg_path=...)            train_cnn.py:50-78
  gram(f)                                             train_cnn.py:101-107
and the loop body / style-setup blocks restated as functions:
  style_grams_single(...)                             train_cnn.py:184-190 ('random'), :199-204 ('average')
  style_grams_smartaverage(...)                       train_cnn.py:224-244
  StyleGramBank                                       train_cnn.py:206-223,316-320 ('cycle': per-painting Grams on device)
  perceptual_step(...)                                train_cnn.py:295-333
  PerceptualTrainer                                   train_cnn.py:247-248,295-334,375 (+ data-parallel allreduce)
All arithmetic runs in libast_b200.so; `nn.MSELoss` on the returned tensors still works and differentiates.
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib
from . import cnn as _cnn
from . import conv_geometry as cg
from . import dp
from . import ops
from .cnn import _ConvParams

CONTENT_WEIGHT = 17  # train_cnn.py:40
STYLE_WEIGHT = 25    # train_cnn.py:41
LR = 0.0024          # train_cnn.py:38
IMAGENET_NEG_MEAN = (-103.939, -116.779, -123.68)  # BGR, train_cnn.py:164

_VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
_TAPS = {3: "relu1_2", 8: "relu2_2", 15: "relu3_3", 22: "relu4_3"}


def _vgg_layout():
    """[(idx, kind, cin, cout)] of torchvision vgg16().features."""
    layers, cin, idx = [], 3, 0
    for v in _VGG_CFG:
        if v == "M":
            layers.append((idx, "pool", cin, cin)); idx += 1
        else:
            layers.append((idx, "conv", cin, v)); layers.append((idx + 1, "relu", v, v)); idx += 2
            cin = v
    return layers


def _vgg_forward(x, module, upto, shift, only_last, need_grad):
    """conv3x3(pad 1)+ReLU / maxpool chain up to `upto` on NHWC buffers.  Returns (taps [(idx, NHWC tensor)], plan, tensor)
    where plan = [(idx, kind, input, output, pool codes)] is what `_vgg_backward` walks."""
    if not x.is_cuda:
        raise RuntimeError("VGG16 kernels run on CUDA only (no CPU fallback)")
    tensor = module._mode() == "fast" and _lib.has_tc_conv()
    x32 = x.detach()
    if x32.dtype not in (torch.float32, torch.uint8):      # uint8 images are widened inside conv1_1's loader
        x32 = x32.to(torch.float32)
    n, _, h, w = x32.shape
    dev = x32.device
    cur = x32.permute(0, 2, 3, 1)            # NCHW tensor described as an (N,H,W,C) view; conv1_1 reads it directly
    taps, plan = [], []
    pooled_next, codes_next = None, None
    packed = module._packed(tensor)
    # fast mode: every activation that only feeds the next convolution (and the ReLU masks of the backward) is stored as
    # fp16 - the 10 mantissa bits a kind::tf32 MMA keeps of an fp32 operand, at half the bytes and twice the MMA rate
    # (kind::f16, fp32 accumulate); the four tap activations stay fp32 (TF32-rounded) for the Gram / MSE kernels
    act16 = torch.float16 if tensor else torch.float32
    packed16 = module._packed_half() if tensor else None
    for idx, kind, cin, cout in module._layout:
        if idx > upto:
            break
        is_tap = (idx + 1) in _TAPS
        odt = torch.float32 if (is_tap or not tensor) else act16
        if kind == "conv" and idx == 0 and tensor:
            # conv1_1 on the tensor cores: fold the 3 horizontal taps (and the mean shift, applied before the
            # zero padding like train_cnn.py:300-301) into a 32-channel fp16 tensor (9 used), then a 3-tap vertical conv
            xr = torch.empty((n, h, w, 32), dtype=act16, device=dev)
            ops.row_im2col(cur, xr, 3, 1, 1, 0, False, shift=shift)
            launches = [cg.Launch(h, w, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
            out = torch.empty((n, h, w, cout), dtype=odt, device=dev)
            wst, stk, bias = module._packed_conv11_stacked(launches[0])
            ops.conv_stacked(xr, wst, stk, out, bias=bias, relu=True)
            plan.append((idx, "conv", cur, out, None))
            cur = out
        elif kind == "conv":
            if cur.dtype == torch.uint8:        # strict mode: the FFMA conv reads fp32
                f32 = torch.empty(cur.shape, dtype=torch.float32, device=dev)
                ops.copy_image(cur, f32)
                cur = f32
            launches = cg.conv_fwd(3, 1, 1, cur.shape[1], cur.shape[2])
            out = torch.empty((n, cur.shape[1], cur.shape[2], cout), dtype=odt, device=dev)
            wp, bias = packed[idx]
            use_tc = tensor and ops.tc_eligible(cur, cout)
            if tensor:
                if not use_tc:
                    raise RuntimeError(f"VGG16 fast mode: conv {idx} ({cin}->{cout}) is not eligible for the tensor-core kernels")
                wp = packed16[idx]
            rnd = tensor and odt == torch.float32        # fp32 tap outputs are rounded to TF32 for the Gram / MSE kernels
            # conv1_2 -> ReLU -> MaxPool2d: the weight-stationary kernel also writes the pooled tensor (saves the pool
            # kernel's read of relu1_2 at B=32) and, when a backward will follow, the 1-byte window codes it needs instead
            # of the activations; when nothing needs the full-resolution relu1_2 (no-grad content branch asking only
            # for its last tap) it is not even stored
            fuse_pool = (use_tc and cin == 64 and cout == 64 and idx + 2 <= upto
                         and cur.shape[1] % 2 == 0 and cur.shape[2] % 2 == 0)
            if fuse_pool:
                hp, wp_ = cur.shape[1] // 2, cur.shape[2] // 2
                pooled_next = torch.empty((n, hp, wp_, cout), dtype=act16, device=dev)
                codes_next = torch.empty((n, hp, wp_, cout), dtype=torch.uint8, device=dev) if need_grad else None
                skip_full = only_last and not need_grad
                ops.conv_gather(cur, wp, launches, out, bias=bias, relu=True, tensor=True, round_tf32=rnd,
                                pooled=pooled_next, pool_only=skip_full, pool_codes=codes_next)
                if skip_full:
                    out = None
            else:
                ops.conv_gather(cur, wp, launches, out, bias=bias, in_shift=shift if idx == 0 else None, relu=True,
                                tensor=use_tc, round_tf32=rnd)
            plan.append((idx, "conv", cur, out, None))
            cur = out
        elif kind == "pool":
            codes = None
            if pooled_next is not None:
                out, codes, pooled_next, codes_next = pooled_next, codes_next, None, None
            else:
                if need_grad and cur.shape[1] % 2 == 0 and cur.shape[2] % 2 == 0 and cur.shape[3] % 8 == 0:
                    codes = torch.empty((n, cur.shape[1] // 2, cur.shape[2] // 2, cur.shape[3]), dtype=torch.uint8, device=dev)
                out = ops.maxpool2_fwd(cur, codes=codes, out_dtype=act16 if tensor else None)
            plan.append((idx, "pool", cur, out, codes))
            cur = out
        if idx in _TAPS and cur is not None:
            taps.append((idx, cur))
    if only_last:
        taps = taps[-1:]
    return taps, plan, tensor


def _vgg_backward(module, plan, tensor, tapg):
    """Data gradient of the chain w.r.t. its input image.  tapg: {tap index: NHWC gradient (fp32 or bf16, contiguous)}."""
    # fast mode: the gradient chain through the frozen VGG runs in bf16 (fp32 accumulation), like the transform
    # net's backward; the forward taps, Grams and losses keep TF32 (parity is stated on those).
    bf16_bwd = tensor
    gdt = torch.bfloat16 if bf16_bwd else torch.float32
    packed = module._packed_dgrad(tensor, bf16_bwd)
    g = None            # gradient w.r.t. the OUTPUT of the current plan entry (already ReLU-masked for convs)
    gx = None
    for pos in reversed(range(len(plan))):
        idx, kind, xin, out, codes = plan[pos]
        if kind == "pool":
            if g is None:
                continue
            # out = pool(xin); xin is the ReLU output of the conv before: route + add its tap grad + mask.  With the
            # forward's window codes the activations are not re-read (1 byte per window instead of 4 x fp32)
            prev_relu_idx = idx - 1
            g = ops.maxpool2_bwd(xin, g, tapg.pop(prev_relu_idx, None), codes=codes)
            continue
        relu_idx = idx + 1
        if relu_idx in tapg:   # tap gradient not yet folded in (only when no pool/conv consumer did it)
            t = tapg.pop(relu_idx)
            m = torch.empty(t.shape, dtype=gdt, device=t.device)
            ops.mask_add(t, g, out, m)       # (tap grad + downstream grad) * (relu out > 0)
            g = m
        if g is None:
            continue
        # g is d/d(relu out) masked == d/d(conv out).  dgrad to the conv input:
        if idx == 0:
            gx = torch.empty((out.shape[0], 3, out.shape[1], out.shape[2]), dtype=torch.float32, device=g.device)
            if tensor and g.dtype == torch.bfloat16:
                # d(image) of conv1_1 as 3 VERTICAL taps whose 9 (of 32) output channels are the partial sums of the
                # 3 horizontal taps x 3 image channels, finished by ast_fold_rows - 12 instead of 36 N=32 MMAs per
                # 128 pixels and no strided 3-channel epilogue:
                #   P[a][u][kx][ci] = sum_{ky,co} g[a-ky+1][u][co] W[co][ci][ky][kx],  gx[a][b][ci] = sum_kx P[a][b-kx+1][kx][ci]
                hh, ww = out.shape[1], out.shape[2]
                part = torch.empty((out.shape[0], hh, ww + 2, 32), dtype=torch.float32, device=g.device)
                taps = [(1 - ky, -1) for ky in range(3)]
                lv = [cg.Launch(hh, ww + 2, 1, 1, 0, 0, taps, [(ky, 0) for ky in range(3)], 0)]
                wst, stk = module._packed_conv11_vdgrad_stacked(g.dtype, lv[0])
                ops.conv_stacked(g, wst, stk, part)
                ops.fold_rows(part, gx.permute(0, 2, 3, 1), 3)
            else:
                launches = cg.conv_dgrad(3, 1, 1, out.shape[1], out.shape[2])
                ops.conv_gather(g, packed[idx], launches, gx.permute(0, 2, 3, 1), tensor=tensor)
            break
        launches = cg.conv_dgrad(3, 1, 1, xin.shape[1], xin.shape[2])
        gin = torch.empty(xin.shape, dtype=gdt, device=g.device)
        # the conv input is either a ReLU output (mask here, add its tap grad) or a pool output (no mask)
        prev_kind = plan[pos - 1][1]
        use_tc = tensor and ops.tc_eligible(g, xin.shape[3])
        if prev_kind == "conv":
            ops.conv_gather(g, packed[idx], launches, gin, add=tapg.pop(idx - 1, None), mask=xin, tensor=use_tc,
                            round_tf32=tensor and not bf16_bwd)
        else:
            ops.conv_gather(g, packed[idx], launches, gin, tensor=use_tc, round_tf32=tensor and not bf16_bwd)
        g = gin
    return gx


class _VGGFunction(torch.autograd.Function):
    """VGG16.forward as an autograd node returning the tapped activations (NCHW views)."""

    @staticmethod
    def forward(ctx, x, module, upto, shift, only_last, *weights):
        taps, plan, tensor = _vgg_forward(x, module, upto, shift, only_last, ctx.needs_input_grad[0])
        ctx.module, ctx.plan, ctx.tensor = module, plan, tensor
        ctx.tap_idx = [i for i, _ in taps]
        return tuple(t.permute(0, 3, 1, 2) for _, t in taps)

    @staticmethod
    def backward(ctx, *gtaps):
        tapg = {}
        for idx, g in zip(ctx.tap_idx, gtaps):
            if g is not None:
                g = g.to(torch.float32).permute(0, 2, 3, 1)
                if not g.is_contiguous():
                    gc = torch.empty(g.shape, dtype=torch.float32, device=g.device)
                    ops.copy_image(g, gc)
                    g = gc
                tapg[idx] = g
        gx = _vgg_backward(ctx.module, ctx.plan, ctx.tensor, tapg)
        ctx.plan = None
        return (gx, None, None, None, None) + tuple(None for _ in ctx.module._weights())


class VGG16(nn.Module, _cnn._Precision):
    """Frozen VGG16 feature extractor returning {'relu1_2','relu2_2','relu3_3','relu4_3'} (train_cnn.py:50-78).

    Input is BGR, 0-255, Caffe-mean-subtracted (train_cnn.py:164,300-301).  `forward(x, shift=...)` can fuse
    that subtraction into conv1_1's loader; `forward(x, upto='relu2_2')` stops early (content branch).
    """

    def __init__(self, just_content=False, vgg_path="models/vgg16-00b39a1b.pth", precision=None):
        super().__init__()
        self.precision = precision
        self._layout = _vgg_layout()
        mods = []
        for idx, kind, cin, cout in self._layout:
            if kind == "conv":
                m = _ConvParams((cout, cin, 3, 3), cin * 9, cout)
                with torch.no_grad():           # torchvision vgg init: kaiming_normal_(fan_out, relu), bias 0
                    m.weight.normal_(0.0, math.sqrt(2.0 / (cout * 9)))
                    m.bias.zero_()
            elif kind == "relu":
                m = nn.ReLU(inplace=True)       # structural placeholders; compute is in libast_b200.so
            else:
                m = nn.MaxPool2d(2, 2)
            mods.append(m)
        self.features = nn.Sequential(*mods)
        if vgg_path is not None:
            # like the reference (train_cnn.py:55) a missing weight file is an error, not a silently random VGG;
            # vgg_path=None is the explicit way to ask for random-init weights (tests, bench)
            if not os.path.exists(vgg_path):
                raise FileNotFoundError(f"VGG16 weights not found at {vgg_path!r} (pass vgg_path=None for random-init weights)")
            self.load_state_dict(torch.load(vgg_path, map_location="cpu"), strict=False)   # train_cnn.py:55
        self.just_content = just_content
        for p in self.features.parameters():    # train_cnn.py:60-61
            p.requires_grad = False
        self._pack_cache = {}

    def _weights(self):
        return [p for p in self.features.parameters()]

    def _cache_key(self, kind, tensor):
        return (kind, tensor, tuple(p._version for p in self.features.parameters()),
                tuple(p.data_ptr() for p in self.features.parameters()))

    def _packed(self, tensor):
        key = self._cache_key("fwd", tensor)
        if self._pack_cache.get("fwd_key") != key:
            out = {}
            launches = cg.conv_fwd(3, 1, 1, 8, 8)
            for idx, kind, cin, cout in self._layout:
                if kind == "conv" and idx <= 21:
                    m = self.features[idx]
                    w = m.weight.detach().float()
                    if idx == 0 and tensor:     # [dy][co][dx*3+c (9 of 16)] for the row-im2col'd conv1_1
                        out[idx] = (ops.pack_weights_ex(w, [0, 3, 6], cout, cout, 16, 9, 3, 27, 1, 9, ops.TF32),
                                    m.bias.detach().float())
                        continue
                    out[idx] = (ops.pack_weights(w, launches, cout, cin, cin * 9, 9, 3, 1,
                                                 ops.TF32 if tensor else torch.float32), m.bias.detach().float())
            self._pack_cache["fwd_key"], self._pack_cache["fwd"] = key, out
        return self._pack_cache["fwd"]

    def _packed_dgrad(self, tensor, bf16=False):
        key = self._cache_key("dgrad", (tensor, bf16))
        if self._pack_cache.get("dgrad_key") != key:
            out = {}
            launches = cg.conv_dgrad(3, 1, 1, 8, 8)
            wdt = torch.bfloat16 if bf16 else (ops.TF32 if tensor else torch.float32)
            for idx, kind, cin, cout in self._layout:
                if kind == "conv" and idx <= 21:
                    w = self.features[idx].weight.detach().float()
                    if idx == 0 and tensor:     # [t][ci (3 of 32)][co]: thin-output dgrad on the tensor cores
                        offs = [u * 3 + v for u, v in cg.all_wtaps(launches)]
                        out[idx] = ops.pack_weights_ex(w, offs, 32, cin, cout, cout, 1, 9, 27, 0, wdt)
                        continue
                    out[idx] = ops.pack_weights(w, launches, cin, cout, 9, cin * 9, 3, 1, wdt)
            self._pack_cache["dgrad_key"], self._pack_cache["dgrad"] = key, out
        return self._pack_cache["dgrad"]

    def _packed_conv11_vdgrad(self, dtype):
        """[ky][d*3 + ci (9 of 32)][co] = W[co][ci][ky][2 - d]: conv1_1's data gradient as a 3-vertical-tap conv with the
        horizontal taps in the output channels (see _VGGFunction.backward); cached like the other packs."""
        key = self._cache_key("c11v", dtype)
        if self._pack_cache.get("c11v_key") != key:
            w = self.features[0].weight.detach().float()                       # (64, 3, 3, 3) = [co][ci][ky][kx]
            wv = w.flip(3).permute(2, 3, 1, 0).reshape(3, 9, w.shape[0])       # [ky][d*3+ci][co]
            pk = torch.zeros((3, 32, w.shape[0]), dtype=torch.float32, device=w.device)
            pk[:, :9, :] = wv
            self._pack_cache["c11v_key"], self._pack_cache["c11v"] = key, pk.to(dtype).contiguous()
        return self._pack_cache["c11v"]

    def _packed_half(self):
        """fp16 copies of the forward packs (TF32-rounded values are exactly representable in fp16 inside its range)."""
        key = self._cache_key("fwd16", True)
        if self._pack_cache.get("fwd16_key") != key:
            self._pack_cache["fwd16_key"] = key
            self._pack_cache["fwd16"] = {i: wp.half() for i, (wp, _) in self._packed(True).items() if i != 0}
        return self._pack_cache["fwd16"]

    def _packed_conv11_stacked(self, launch):
        """conv1_1's 3-vertical-tap filter stacked for ast_conv_stacked: two interleaved output rows x 64 channels fill the
        128 TMEM lanes (4 virtual taps instead of 2 x 3 N=64 MMAs per pixel pair)."""
        stk = cg.stack_rows(launch, 2)
        key = self._cache_key("c11s", True)
        if self._pack_cache.get("c11s_key") != key:
            wp, bias = self._packed(True)[0]                                    # [dy][64][16], TF32-rounded
            wp32 = torch.zeros((wp.shape[0], 64, 32), dtype=torch.float16, device=wp.device)   # fp16, K padded to 64 bytes
            wp32[:, :, :16] = wp.half()
            self._pack_cache["c11s_key"] = key
            self._pack_cache["c11s"] = (ops.stack_filter(lambda pos: wp32[pos[0]], stk, 64, 32, torch.float16, wp.device), bias)
        wst, bias = self._pack_cache["c11s"]
        return wst, stk, bias

    def _packed_conv11_vdgrad_stacked(self, dtype, launch):
        """The vertical-tap data-gradient filter of conv1_1 ([ky][32][64]) stacked over four interleaved output rows."""
        stk = cg.stack_rows(launch, 4)
        key = self._cache_key("c11vs", dtype)
        if self._pack_cache.get("c11vs_key") != key:
            pk = self._packed_conv11_vdgrad(dtype)
            self._pack_cache["c11vs_key"] = key
            self._pack_cache["c11vs"] = ops.stack_filter(lambda pos: pk[pos[0]], stk, 32, pk.shape[2], dtype, pk.device)
        return self._pack_cache["c11vs"], stk

    def forward(self, x, shift=None, upto=None, only_last=False):
        if self.just_content or upto == "relu2_2":
            last = 8
        elif upto is None or upto == "relu4_3":
            last = 22
        else:
            last = {v: k for k, v in _TAPS.items()}[upto]
        if x.dim() == 3:
            x = x.unsqueeze(0)
        outs = _VGGFunction.apply(x, self, last, shift, bool(only_last or self.just_content), *self._weights())
        if self.just_content:
            return outs[-1]                                        # train_cnn.py:64-68
        names = [v for k, v in _TAPS.items() if k <= last]
        if only_last:                                              # only the deepest requested tap is materialised
            names = names[-1:]
        return dict(zip(names, outs))                              # insertion order relu1_2..relu4_3 (:70-77)


class _GramFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, fast):
        b, c, h, w = f.shape
        fv = f.detach()
        if fv.dtype not in (torch.float32, torch.bfloat16):
            fv = fv.float()
        ctx.save_for_backward(fv)
        ctx.fast = fast
        xv = fv.permute(0, 2, 3, 1)
        return ops.gram(xv, 1.0 / (c * h * w), tensor=fast and b > 0 and ops.tc_contract_eligible(xv, xv))

    @staticmethod
    def backward(ctx, dg):
        (fv,) = ctx.saved_tensors
        b, c, h, w = fv.shape
        # dF = (dG + dG^T) F / (CHW): a 1x1 gather-conv with per-image C x C weights
        d = ((dg + dg.transpose(1, 2)) * (1.0 / (c * h * w))).to(fv.dtype).contiguous()
        x = fv.permute(0, 2, 3, 1)
        out = torch.empty((b, h, w, c), dtype=torch.float32, device=fv.device)
        launches = cg.conv_fwd(1, 1, 0, h, w)
        fast = ctx.fast
        ops.conv_gather(x, d.view(b, 1, c, c), launches, out, w_img_stride=c * c,
                        tensor=fast and x.is_contiguous() and ops.tc_eligible(x, c), round_tf32=fast)
        return out.permute(0, 3, 1, 2), None


def gram(f, precision=None):
    """G = F F^T / (C H W) for F = f.view(b, c, h*w) (train_cnn.py:103-107). f: [B,C,H,W] -> [B,C,C]."""
    if not f.is_cuda:
        raise RuntimeError("gram() runs on CUDA only (no CPU fallback)")
    mode = precision or _cnn.get_default_precision()
    return _GramFunction.apply(f, mode == "fast")


_neg_mean_cache = {}


def neg_mean(device):
    """(-103.939, -116.779, -123.68) on `device`, created once (no host-to-device copy inside a captured step)."""
    key = str(device)
    if key not in _neg_mean_cache:
        _neg_mean_cache[key] = torch.tensor(IMAGENET_NEG_MEAN, dtype=torch.float32, device=device)
    return _neg_mean_cache[key]


def style_grams_single(vgg, style_tensor, batch_size):
    """'random' / 'average' style setup (train_cnn.py:184-190): one (3,H,W) image -> 4 x [B,C,C].

    The reference expands the image to the batch and runs B identical VGG passes; the Grams are identical
    per row, so one pass is computed and the result expanded (same values, 1/B of the work).
    """
    with torch.no_grad():
        feats = vgg(style_tensor.float().unsqueeze(0) if style_tensor.dim() == 3 else style_tensor.float(),
                    shift=neg_mean(style_tensor.device))
        # B identical rows in the reference: a stride-0 expand (no copies); the fused loss kernel reads one C x C target
        return {k: gram(v).expand(batch_size, -1, -1) for k, v in feats.items()}


def style_grams_smartaverage(vgg, paintings, batch_size, mode="reference", group=None, chunk=16):
    """'smartaverage' artist style (train_cnn.py:224-244).

    mode='reference': sum the VGG features over the artist's paintings, divide by the count, ONE Gram of the
    mean feature (exactly the reference).  mode='mean_gram': mean of per-painting Grams (north-star
    wording).  With `group` (torch.distributed), each rank passes ITS shard of the paintings and the sums are
    all-reduced (NCCL) before the division; `paintings` is a list of (3,H,W) tensors or a [P,3,H,W] tensor.
    The reference pushes one painting (expanded to B identical copies) through the VGG at a time; here `chunk` DIFFERENT
    paintings share one VGG pass and the feature (or Gram) sum over the chunk is taken by the accumulate kernel - same
    sums, far fewer and fuller launches.
    """
    acc, count = None, 0
    with torch.no_grad():
        plist = list(paintings) if not torch.is_tensor(paintings) else list(paintings.unbind(0))
        for i0 in range(0, len(plist), max(1, chunk)):
            part = plist[i0:i0 + max(1, chunk)]
            if any(p.shape != part[0].shape for p in part):          # ragged sizes: one painting per pass
                groups = [[p] for p in part]
            else:
                groups = [part]
            for grp in groups:
                x = torch.stack([p.float() if p.dtype != torch.uint8 else p for p in grp])
                feats = vgg(x, shift=neg_mean(x.device))
                if mode == "reference":
                    cur = {k: v.permute(0, 2, 3, 1) for k, v in feats.items()}       # NHWC views of the taps
                else:
                    cur = {k: gram(v).unsqueeze(1) for k, v in feats.items()}        # [k, C, C] as images [k, 1, C, C]
                if acc is None:
                    acc = {k: torch.zeros((1,) + tuple(v.shape[1:]), dtype=torch.float32, device=v.device) for k, v in cur.items()}
                for k, v in cur.items():
                    ops.accumulate(v, acc[k])                       # train_cnn.py:239 in-place sum (over the chunk's batch)
                count += len(grp)
        total = torch.tensor([float(count)], device=next(iter(acc.values())).device)
        if group is not None:                                        # no group = every painting is local, no exchange
            dp.allreduce_sums(list(acc.values()) + [total], group)  # C2: one exchange per artist
        length = float(total.item())
        out = {}
        for k, v in acc.items():
            if mode == "reference":
                # gram(sum/len) == gram(sum)/len^2: folds train_cnn.py:242-243's divide into a C x C scale
                g = gram(v.permute(0, 3, 1, 2)) / (length * length)
            else:
                g = v[0] / length                                   # [1, C, C]
            out[k] = g.expand(batch_size, -1, -1)
        return out


class _MSEFunction(torch.autograd.Function):
    """weight * mean((a-b)^2) with the gradient w.r.t. `a` produced in the same pass (nn.MSELoss, train_cnn.py:249,307)."""

    @staticmethod
    def forward(ctx, a, b, weight=1.0):
        av = a.detach().permute(0, 2, 3, 1)
        bv = b.detach().permute(0, 2, 3, 1)
        loss = torch.zeros(1, dtype=torch.float32, device=a.device)
        grad = torch.empty(av.shape, dtype=torch.float32, device=a.device) if a.requires_grad else None
        numel = a.numel()
        ops.mse(av, bv, loss, float(weight) / numel, grad, 2.0 * float(weight) / numel)
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None
        return (ctx.grad * g).permute(0, 3, 1, 2), None, None


def mse_loss(a, b, weight=1.0):
    """Fused (weighted) MSE forward+gradient for two [B,C,H,W] tensors (b is treated as a constant)."""
    return _MSEFunction.apply(a, b, weight)


def _loss_forward(names, feats, targets, content_feat, content_weight, style_weight, fast, need_grad):
    """Loss side of train_cnn.py:307-329 over the VGG taps of the generated batch.  feats: NCHW-shaped tensors (any strides).

    Per tap, ONE kernel (ast_gram_mse) computes the upper triangle of the Gram on the tensor cores and - in its finishing
    step - the style-MSE contribution and D = 4 w_s (G - S) / (B C^2 CHW), the symmetric per-image 1x1 weights of the Gram
    backward dF = D F.  The content term on relu2_2 (train_cnn.py:307-308) is one MSE pass producing
    loss and gradient.  Returns (content loss 1-elem fp32, style loss 1-elem fp64, [Gram], state for `_loss_backward`).
    """
    dev = feats[0].device
    b = feats[0].shape[0]
    n_taps = len(feats)
    sizes = [b * f.shape[1] * f.shape[1] for f in feats]
    cnt = b * _lib.GRAM_COUNTERS_PER_IMAGE
    # one zero fill: [style loss (one fp64), content loss, pad | Gram 0..n | ticket counters 0..n (int32 views)]
    zeros = torch.zeros(4 + sum(sizes) + n_taps * cnt, dtype=torch.float32, device=dev)
    style_acc, content_acc = zeros[:2].view(torch.float64), zeros[2:3]
    off = 4
    grams, dmats, saved = [], [], []
    content_grad = None
    for i, f in enumerate(feats):
        _, c, h, w = f.shape
        fv = f.detach()
        if fv.dtype not in (torch.float32, torch.bfloat16):
            fv = fv.float()
        xv = fv.permute(0, 2, 3, 1)
        g = zeros[off:off + sizes[i]].view(b, c, c)
        off += sizes[i]
        counters = zeros[4 + sum(sizes) + i * cnt:4 + sum(sizes) + (i + 1) * cnt].view(torch.int32)
        d = torch.empty((b, c, c), dtype=torch.float32, device=dev) if need_grad else None
        tgt = targets[i].detach()
        if tgt.dtype != torch.float32:
            tgt = tgt.float()
        ops.gram_mse(xv, tgt, g, counters, loss=style_acc, loss_scale=float(style_weight) / (b * c * c), d=d,
                     d_scale=4.0 * float(style_weight) / (float(b) * c * c * c * h * w),
                     tensor=fast and b > 0 and ops.tc_contract_eligible(xv, xv))
        if names[i] == "relu2_2" and content_feat is not None:                       # train_cnn.py:307-308
            numel = f.numel()
            content_grad = torch.empty(xv.shape, dtype=torch.float32, device=dev) if need_grad else None
            ops.mse(xv, content_feat.detach().permute(0, 2, 3, 1), content_acc, float(content_weight) / numel,
                    content_grad, 2.0 * float(content_weight) / numel)
        grams.append(g)
        dmats.append(d)
        saved.append(fv)
    return content_acc, style_acc, grams, (names, saved, dmats, content_grad, fast)


def _loss_backward(state, g_content, g_style, unit_grad, out_dtype):
    """dF per tap as NHWC tensors of `out_dtype`: a 1x1 gather-conv with per-image C x C weights D (+ the content gradient
    in the epilogue of relu2_2's).  unit_grad: the upstream gradients of both losses are exactly 1 - nothing is rescaled."""
    names, feats, dmats, content_grad, fast = state
    outs = []
    for i, fv in enumerate(feats):
        b, c, h, w = fv.shape
        x = fv.permute(0, 2, 3, 1)
        d = dmats[i]
        cgrad = content_grad if names[i] == "relu2_2" else None
        if not unit_grad:
            d = d * g_style
            cgrad = None if cgrad is None else cgrad * g_content
        if d.dtype != fv.dtype:
            d = d.to(fv.dtype)
        out = torch.empty((b, h, w, c), dtype=out_dtype, device=fv.device)
        ops.conv_gather(x, d.view(b, 1, c, c), cg.conv_fwd(1, 1, 0, h, w), out, add=cgrad, w_img_stride=c * c,
                        tensor=fast and x.is_contiguous() and ops.tc_eligible(x, c), round_tf32=fast and out_dtype == torch.float32)
        outs.append(out)
    return outs


class _PerceptualLossFunction(torch.autograd.Function):
    """`_loss_forward` / `_loss_backward` as ONE autograd node over externally computed VGG taps (the drop-in call sites:
    `vgg(generated)` then the losses).  Returns (content_loss, style_loss, grams...) - the Grams are for inspection (no grad)."""

    @staticmethod
    def forward(ctx, content_feat, content_weight, style_weight, fast, unit_grad, n_taps, *rest):
        feats, targets, names = rest[:n_taps], rest[n_taps:2 * n_taps], rest[2 * n_taps]
        need_grad = any(ctx.needs_input_grad[6:6 + n_taps])
        content_acc, style_acc, grams, ctx.state = _loss_forward(names, feats, targets, content_feat, content_weight,
                                                                 style_weight, fast, need_grad)
        ctx.unit_grad, ctx.n_taps = unit_grad, n_taps
        ctx.mark_non_differentiable(*grams)
        return (content_acc[0], style_acc[0].to(torch.float32), *grams)

    @staticmethod
    def backward(ctx, g_content, g_style, *_):
        outs = _loss_backward(ctx.state, g_content, g_style, ctx.unit_grad, torch.float32)
        ctx.state = None
        return (None, None, None, None, None, None, *[o.permute(0, 3, 1, 2) for o in outs], *([None] * (ctx.n_taps + 1)))


class _VGGPerceptualFunction(torch.autograd.Function):
    """VGG16(generated) + the whole loss side as ONE autograd node: the training path of `perceptual_step`.

    Nothing between the taps and the losses has to be an autograd tensor, so the tap gradients stay in the kernels' own
    format: bf16 NHWC (half the bytes of the fp32 NCHW gradients autograd would demand), consumed directly by the epilogues
    of the VGG data-gradient convolutions and the pooling backward (which reads the forward's 1-byte window codes, not the
    activations).  Same arithmetic as `vgg(x)` followed by `perceptual_losses(...)`."""

    @staticmethod
    def forward(ctx, generated, module, shift, content_feat, content_weight, style_weight, unit_grad, names, *targets):
        need_grad = ctx.needs_input_grad[0]
        last = max(k for k, v in _TAPS.items() if v in names)
        taps, plan, tensor = _vgg_forward(generated, module, last, shift, False, need_grad)
        feats = [t.permute(0, 3, 1, 2) for idx, t in taps if _TAPS[idx] in names]
        tap_names = tuple(_TAPS[idx] for idx, _ in taps if _TAPS[idx] in names)
        tgt = dict(zip(names, targets))
        content_acc, style_acc, grams, state = _loss_forward(tap_names, feats, [tgt[k] for k in tap_names], content_feat,
                                                             content_weight, style_weight, module._mode() == "fast", need_grad)
        ctx.module, ctx.plan, ctx.tensor, ctx.state = module, plan, tensor, state
        ctx.tap_idx = [idx for idx, _ in taps if _TAPS[idx] in names]
        ctx.unit_grad = unit_grad
        ctx.mark_non_differentiable(*grams)
        return (content_acc[0], style_acc[0].to(torch.float32), *grams)

    @staticmethod
    def backward(ctx, g_content, g_style, *_):
        gdt = torch.bfloat16 if ctx.tensor else torch.float32
        outs = _loss_backward(ctx.state, g_content, g_style, ctx.unit_grad, gdt)
        gx = _vgg_backward(ctx.module, ctx.plan, ctx.tensor, dict(zip(ctx.tap_idx, outs)))
        ctx.plan = ctx.state = None
        return (gx, None, None, None, None, None, None, None, *([None] * len(ctx.tap_idx)))


def perceptual_losses(gen_feats, content_feat, style_gram, content_weight=CONTENT_WEIGHT, style_weight=STYLE_WEIGHT,
                      fast=None, unit_grad=False):
    """(content_loss, style_loss, {tap: Gram}) of train_cnn.py:307-325 for the generated batch's VGG taps."""
    names = tuple(gen_feats.keys())
    feats = [gen_feats[k] for k in names]
    targets = [style_gram[k] for k in names]
    if fast is None:
        fast = _cnn.get_default_precision() == "fast"
    out = _PerceptualLossFunction.apply(content_feat, content_weight, style_weight, bool(fast), bool(unit_grad), len(names),
                                        *feats, *targets, names)
    return out[0], out[1], dict(zip(names, out[2:]))


def perceptual_step(transfer, vgg, content_batch, style_gram, content_weight=CONTENT_WEIGHT,
                    style_weight=STYLE_WEIGHT, backward=True):
    """The loop body of train_cnn.py:295-333 (methods 'random'/'average'/'cycle'/'smartaverage'), without the optimizer.

    content_batch: [B,3,H,W] BGR 0-255, fp32 or uint8.  Returns (content_loss, style_loss, total_loss) as 0-d device
    tensors (no host sync).  Differences from the reference that do not change results: the mean shift is fused into
    conv1_1's loader, and the content branch stops at relu2_2 (the reference computes relu3_3/relu4_3 and discards them).
    """
    shift = neg_mean(content_batch.device)
    generated = transfer(content_batch)                                        # :299
    with torch.no_grad():
        content_feat = vgg(content_batch, shift=shift, upto="relu2_2", only_last=True)["relu2_2"]   # :300
    names = tuple(k for k in _TAPS.values() if k in style_gram)
    out = _VGGPerceptualFunction.apply(generated, vgg, shift, content_feat, content_weight, style_weight, bool(backward),
                                       names, *[style_gram[k] for k in names])                   # :301, :307-325
    content_loss, style_loss = out[0], out[1]
    total = content_loss + style_loss                                          # :329
    if backward:
        total.backward()                                                       # :333
    return content_loss.detach(), style_loss.detach(), total.detach()


class StyleGramBank:
    """'cycle' style method (train_cnn.py:206-223, 316-320): the Grams of every painting of the artist stay resident on
    the device ([P,C,C] per tap, 1.4 MB per painting) instead of being shuttled to the CPU and back every step
    (`.cpu()` at :218, `.to(device)` at :323); step t uses painting t % P.  Each painting's Gram is shared by the batch
    (the reference stores B identical rows), so targets are [C,C] views."""

    def __init__(self, vgg, paintings):
        per = [style_grams_single(vgg, p, 1) for p in paintings]
        self.keys = list(per[0].keys())
        self.bank = {k: torch.cat([g[k] for g in per], dim=0).contiguous() for k in self.keys}    # [P, C, C]
        self.length = len(per)

    def __len__(self):
        return self.length

    def target(self, index):
        i = index % self.length                                               # :317
        return {k: v[i] for k, v in self.bank.items()}


class _Prefetched:
    """Handle returned by PerceptualTrainer.prefetch(): a device staging buffer and the event of its H2D copy."""
    __slots__ = ("tensor", "ready", "slot")

    def __init__(self, tensor, ready, slot):
        self.tensor, self.ready, self.slot = tensor, ready, slot


class PerceptualTrainer:
    """Optimizer side of train() (train_cnn.py:247-248,295,334,375) plus data-parallel gradient averaging.

    optimizer="fused" (default): the TransformerNet gradients land in one flat arena (arena.py) that is all-reduced as it
    is (one NCCL call) and consumed by ONE Adam(L2) kernel which also refreshes the packed bf16 weights of
    the next step; lr and the step count live in device memory, so StepLR (end_epoch) also acts on CUDA-graph replays.
    optimizer="torch": torch.optim.Adam + StepLR on the parameters' .grad (the reference's objects, :247-248).
    One process per GPU; with torch.distributed initialised, rank 0's parameters are broadcast at construction so that
    replicas start identical, and every step averages the gradients.
    """

    def __init__(self, transfer, vgg, style_gram, lr=LR, weight_decay=1e-4, num_epochs=200, num_steps=2,
                 content_weight=CONTENT_WEIGHT, style_weight=STYLE_WEIGHT, group=None, cuda_graph=False,
                 optimizer="fused"):
        self.transfer, self.vgg, self.style_gram = transfer, vgg, style_gram
        self.content_weight, self.style_weight = content_weight, style_weight
        self.params = [p for p in transfer.parameters()]
        on_cuda = self.params[0].is_cuda
        self.group = group
        self.world, self.rank = dp.world(group)
        if self.world > 1:
            dp.broadcast_parameters(self.params, group)          # replicas must start identical (one flat call)
        self.base_lr, self.lr, self.gamma, self.epoch = float(lr), float(lr), 0.5, 0
        self.step_size = max(1, num_epochs // num_steps)                                           # :248
        self.cuda_graph = bool(cuda_graph and on_cuda)
        self.fused = optimizer == "fused" and on_cuda
        if optimizer not in ("fused", "torch"):
            raise ValueError("optimizer must be 'fused' or 'torch'")
        self._flat = None
        if self.fused:
            self.arena = transfer._arena_for(self.params[0].device)
            self.arena.enable_optimizer(lr, weight_decay=weight_decay)
            self.gbuf = self.arena.new_grad_buffer()
            self.arena.grad_sink = self.gbuf
            for p, g in zip(self.arena.params(), self.arena.grad_views(self.gbuf)):
                p.grad = g                                       # views of the arena (strided for conv weights)
            self.optimizer = self.scheduler = None
        else:
            # lr as a device tensor under graph capture: torch's fused Adam then reads it at replay time, so StepLR's
            # in-place update is seen by the captured graph (a Python float would be frozen into it)
            lr_arg = torch.tensor(lr, dtype=torch.float32, device=self.params[0].device) if self.cuda_graph else lr
            self.optimizer = torch.optim.Adam(self.params, lr=lr_arg, weight_decay=weight_decay, fused=bool(on_cuda),
                                              capturable=self.cuda_graph)                         # :247
            self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=self.step_size, gamma=0.5)
        # cuda_graph=True: after 3 eager warm-up steps the whole step (fwd, bwd, all-reduce, Adam: ~300 launches) is
        # captured once per input shape and replayed, removing host launch overhead and inter-kernel gaps.
        self._graph, self._static_in, self._static_losses, self._eager_steps = None, None, None, 0
        self._static_style = None
        self._copy_stream = None

    # ---- host -> device input pipeline --------------------------------------------------------------------------
    def prefetch(self, host_batch):
        """Start copying a (pinned) host batch to the device on a side stream and return a handle that `step()` accepts.

        host_batch: [B,3,H,W] fp32 or uint8 (uint8 moves 4x fewer bytes over PCIe; values are widened in the first
        layers' loaders).  Called for batch i+1 before `step(batch i)` is waited on, the copy overlaps the compute of
        step i.  Two device staging buffers alternate; a buffer is only overwritten after the step that consumed it has
        copied it into the step's input.
        """
        dev = self.params[0].device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage, self._stage_free, self._stage_i = [None, None], [None, None], 0
        i = self._stage_i
        self._stage_i ^= 1
        cs = self._copy_stream
        if self._stage[i] is None or self._stage[i].shape != host_batch.shape or self._stage[i].dtype != host_batch.dtype:
            # allocate ON the copy stream: the caching allocator may hand out a block whose previous user still has
            # kernels queued on the compute stream, so order the copy stream behind it first
            cs.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cs):
                self._stage[i] = torch.empty(host_batch.shape, dtype=host_batch.dtype, device=dev)
            self._stage_free[i] = None
        with torch.cuda.stream(cs):
            if self._stage_free[i] is not None:
                cs.wait_event(self._stage_free[i])
            self._stage[i].copy_(host_batch, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        return _Prefetched(self._stage[i], ready, i)

    def _consume(self, batch):
        """Tensor (host or device) or prefetch handle -> device tensor usable on the current stream."""
        if isinstance(batch, _Prefetched):
            torch.cuda.current_stream().wait_event(batch.ready)
            return batch.tensor, batch.slot
        return batch, None

    def _release(self, slot):
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._stage_free[slot] = ev

    def _allreduce_grads(self):
        if self.world == 1:
            return
        if self.fused:
            dp.allreduce_mean_flat(self.gbuf, self.group)         # the arena IS the bucket
        else:
            if self._flat is None:
                self._flat = dp.GradBucket()
            self._flat.allreduce_mean(self.params, self.group)

    def _eager_step(self, content_batch, style_gram):
        if self.fused:
            self.gbuf.zero_()                                                                      # :295
        else:
            self.optimizer.zero_grad(set_to_none=True)
        losses = perceptual_step(self.transfer, self.vgg, content_batch, style_gram,
                                 self.content_weight, self.style_weight, backward=True)
        self._allreduce_grads()
        if self.fused:
            self.arena.adam_step(self.gbuf)                                                        # :334
        else:
            self.optimizer.step()
        return losses

    def _capture(self, content_batch, style_gram):
        self._static_in = content_batch.clone()
        # static copies of the style targets: 'cycle' switches the target per step by copying 1.4 MB into them
        self._static_style = {k: (v[0:1].clone().expand_as(v) if v.dim() == 3 and v.stride(0) == 0 else v.clone())
                              for k, v in style_gram.items()}
        self._graph = torch.cuda.CUDAGraph()
        if not self.fused:
            self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph):
            self._static_losses = self._eager_step(self._static_in, self._static_style)

    def step(self, content_batch, style_gram=None):
        """One optimisation step on `content_batch`: a device tensor, a host tensor, or a handle from `prefetch()`.
        style_gram: optional per-step style targets ('cycle': `bank.target(t)`); default: the trainer's own."""
        content_batch, slot = self._consume(content_batch)
        sg = self.style_gram if style_gram is None else style_gram
        if not self.cuda_graph:
            if not content_batch.is_cuda:
                content_batch = content_batch.to(self.params[0].device, non_blocking=True)
            losses = self._eager_step(content_batch, sg)
            self._release(slot)
            return losses
        if (self._graph is not None and self._static_in.shape == content_batch.shape
                and self._static_in.dtype == content_batch.dtype
                and all(self._static_style[k].shape == sg[k].shape for k in sg)):
            self._static_in.copy_(content_batch, non_blocking=True)
            if style_gram is not None:
                for k, v in sg.items():
                    dst = self._static_style[k]
                    if dst.dim() == 3 and dst.stride(0) == 0:          # one shared target behind a stride-0 expand
                        dst[0].copy_(v[0] if v.dim() == 3 else v, non_blocking=True)
                    else:
                        dst.copy_(v, non_blocking=True)
            self._release(slot)
            self._graph.replay()
            return self._static_losses
        if not content_batch.is_cuda:
            content_batch = content_batch.to(self.params[0].device, non_blocking=True)
        if self._eager_steps < 3:                       # warm up caches (packed weights, tap tables, allocator)
            self._eager_steps += 1
            losses = self._eager_step(content_batch, sg)
            self._release(slot)
            return losses
        self._graph = None
        self._capture(content_batch, sg)                # capture does not execute: run the graph once for this batch
        self._release(slot)
        self._graph.replay()
        return self._static_losses

    def end_epoch(self):
        """StepLR(step_size=num_epochs // num_steps, gamma=0.5).step() (train_cnn.py:248,375)."""
        self.epoch += 1
        if self.fused:
            self.lr = self.base_lr * self.gamma ** (self.epoch // self.step_size)
            self.arena.set_lr(self.lr)                  # device scalar: graph replays pick it up
        else:
            self.scheduler.step()
            self.lr = float(self.scheduler.get_last_lr()[0])

    def close(self):
        """Drop the captured graph (and its NCCL work) BEFORE the process group is destroyed: tearing down a communicator
        that a live CUDA graph still references blocked interpreter exit on the B200 boxes."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._static_losses = None
        if self.fused:
            self.arena.grad_sink = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
