"""B200-native mirror of the loss side of `/root/reference/train_cnn.py`.

Drop-in pieces (same names / signatures / return types as the reference):
  VGG16(just_content=False, vgg_path=...)            train_cnn.py:50-78
  gram(f)                                             train_cnn.py:101-107
and the loop body / style-setup blocks restated as functions:
  style_grams_single(...)                             train_cnn.py:184-190 ('random'), :199-204 ('average')
  style_grams_smartaverage(...)                       train_cnn.py:224-244
  perceptual_step(...)                                train_cnn.py:295-333
  PerceptualTrainer                                   train_cnn.py:247-248,295-334 (+ data-parallel allreduce)
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


def _conv11_fold():
    return os.environ.get("AST_CONV11_FOLD", "1") == "1"


def _fuse_pool():
    return os.environ.get("AST_FUSE_POOL", "1") == "1"


def _vgg_bwd_bf16():
    return os.environ.get("AST_VGG_BWD", "bf16") != "tf32"


class _VGGFunction(torch.autograd.Function):
    """conv3x3(pad 1)+ReLU / maxpool chain up to `upto`, returning the tapped activations (NCHW views)."""

    @staticmethod
    def forward(ctx, x, module, upto, shift, only_last, *weights):
        if not x.is_cuda:
            raise RuntimeError("VGG16 kernels run on CUDA only (no CPU fallback)")
        tensor = module._mode() == "fast" and _lib.has_tc_conv()
        x32 = x.detach().to(torch.float32)
        n, _, h, w = x32.shape
        dev = x32.device
        cur = x32.permute(0, 2, 3, 1)            # NCHW tensor described as an (N,H,W,C) view; conv1_1 reads it directly
        acts, taps, plan = {}, [], []
        pooled_next = None
        packed = module._packed(tensor)
        for idx, kind, cin, cout in module._layout:
            if idx > upto:
                break
            if kind == "conv" and idx == 0 and tensor:
                # conv1_1 on the tensor cores: fold the 3 horizontal taps (and the mean shift, applied before the
                # zero padding like train_cnn.py:300-301) into a 16-channel TF32 tensor, then a 3-tap vertical conv
                xr = torch.empty((n, h, w, 16), dtype=torch.float32, device=dev)
                ops.row_im2col(cur, xr, 3, 1, 1, 0, False, shift=shift, round_tf32=True)
                launches = [cg.Launch(h, w, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
                out = torch.empty((n, h, w, cout), dtype=torch.float32, device=dev)
                wp, bias = packed[idx]
                ops.conv_gather(xr, wp, launches, out, bias=bias, relu=True, tensor=True, round_tf32=True)
                plan.append((idx, "conv", cur, out))
                cur = out
            elif kind == "conv":
                launches = cg.conv_fwd(3, 1, 1, cur.shape[1], cur.shape[2])
                out = torch.empty((n, cur.shape[1], cur.shape[2], cout), dtype=torch.float32, device=dev)
                wp, bias = packed[idx]
                use_tc = tensor and ops.tc_eligible(cur, cout)
                # conv1_2 -> ReLU -> MaxPool2d: the weight-stationary kernel also writes the pooled tensor (saves the pool
                # kernel's 537 MB read at B=32); when nothing needs the full-resolution relu1_2 (no-grad content branch
                # asking only for its last tap) it is not even stored
                fuse_pool = (use_tc and _fuse_pool() and cin == 64 and cout == 64 and idx + 2 <= upto
                             and cur.shape[1] % 2 == 0 and cur.shape[2] % 2 == 0)
                if fuse_pool:
                    pooled_next = torch.empty((n, cur.shape[1] // 2, cur.shape[2] // 2, cout), dtype=torch.float32, device=dev)
                    skip_full = only_last and not ctx.needs_input_grad[0]
                    ops.conv_gather(cur, wp, launches, out, bias=bias, relu=True, tensor=True, round_tf32=True,
                                    pooled=pooled_next, pool_only=skip_full)
                    if skip_full:
                        out = None
                else:
                    ops.conv_gather(cur, wp, launches, out, bias=bias, in_shift=shift if idx == 0 else None, relu=True,
                                    tensor=use_tc, round_tf32=tensor)
                plan.append((idx, "conv", cur, out))
                cur = out
            elif kind == "pool":
                if pooled_next is not None:
                    out, pooled_next = pooled_next, None
                else:
                    out = ops.maxpool2_fwd(cur)
                plan.append((idx, "pool", cur, out))
                cur = out
            if idx in _TAPS and cur is not None:
                taps.append((idx, cur))
        ctx.module, ctx.plan, ctx.tensor = module, plan, tensor
        ctx.tap_idx = [i for i, _ in taps]
        if only_last:
            taps = taps[-1:]
            ctx.tap_idx = ctx.tap_idx[-1:]
        outs = tuple(t.permute(0, 3, 1, 2) for _, t in taps)
        return outs

    @staticmethod
    def backward(ctx, *gtaps):
        module, plan, tensor = ctx.module, ctx.plan, ctx.tensor
        tapg = {}
        for idx, g in zip(ctx.tap_idx, gtaps):
            if g is not None:
                g = g.to(torch.float32).permute(0, 2, 3, 1)
                if not g.is_contiguous():
                    gc = torch.empty(g.shape, dtype=torch.float32, device=g.device)
                    ops.copy_image(g, gc)
                    g = gc
                tapg[idx] = g
        # fast mode: the gradient chain through the frozen VGG runs in bf16 (fp32 accumulation), like the transform
        # net's backward; the forward taps, Grams and losses keep TF32 (parity is stated on those).  AST_VGG_BWD=tf32
        # restores a TF32 backward.
        bf16_bwd = tensor and _vgg_bwd_bf16()
        gdt = torch.bfloat16 if bf16_bwd else torch.float32
        packed = module._packed_dgrad(tensor, bf16_bwd)
        g = None            # gradient w.r.t. the OUTPUT of the current plan entry (already ReLU-masked for convs)
        gx = None
        for pos in reversed(range(len(plan))):
            idx, kind, xin, out = plan[pos]
            if kind == "pool":
                if g is None:
                    continue
                # out = pool(xin); xin is the ReLU output of the conv before: route + add its tap grad + mask
                prev_relu_idx = idx - 1
                g = ops.maxpool2_bwd(xin, g, tapg.pop(prev_relu_idx, None))
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
                if tensor and _conv11_fold() and g.dtype == torch.bfloat16:
                    # d(image) of conv1_1 as 3 VERTICAL taps whose 9 (of 32) output channels are the partial sums of the
                    # 3 horizontal taps x 3 image channels, finished by ast_fold_rows - 12 instead of 36 N=32 MMAs per
                    # 128 pixels and no strided 3-channel epilogue:
                    #   P[a][u][kx][ci] = sum_{ky,co} g[a-ky+1][u][co] W[co][ci][ky][kx],  gx[a][b][ci] = sum_kx P[a][b-kx+1][kx][ci]
                    hh, ww = out.shape[1], out.shape[2]
                    part = torch.empty((out.shape[0], hh, ww + 2, 32), dtype=torch.float32, device=g.device)
                    taps = [(1 - ky, -1) for ky in range(3)]
                    lv = [cg.Launch(hh, ww + 2, 1, 1, 0, 0, taps, [(ky, 0) for ky in range(3)], 0)]
                    ops.conv_gather(g, module._packed_conv11_vdgrad(g.dtype), lv, part, tensor=True)
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
        ctx.plan = None
        return (gx, None, None, None, None) + tuple(None for _ in module._weights())


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
        if vgg_path is not None and os.path.exists(vgg_path):
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
        return {k: gram(v).expand(batch_size, -1, -1).contiguous() for k, v in feats.items()}


def style_grams_smartaverage(vgg, paintings, batch_size, mode="reference", group=None):
    """'smartaverage' artist style (train_cnn.py:224-244).

    mode='reference': sum the VGG features over the artist's paintings, divide by the count, ONE Gram of the
    mean feature (exactly the reference).  mode='mean_gram': mean of per-painting Grams (north-star
    wording).  With `group` (torch.distributed), each rank passes ITS shard of the paintings and the sums are
    all-reduced (NCCL) before the division; `paintings` is a list of (3,H,W) tensors or a [P,3,H,W] tensor.
    """
    acc, count = None, 0
    shift = None
    with torch.no_grad():
        for p in paintings:
            shift = neg_mean(p.device) if shift is None else shift
            feats = vgg(p.float().unsqueeze(0), shift=shift)
            cur = {}
            for k, v in feats.items():
                cur[k] = v.permute(0, 2, 3, 1) if mode == "reference" else gram(v)
            if acc is None:
                acc = {k: torch.zeros(v.shape, dtype=torch.float32, device=v.device) for k, v in cur.items()}
            for k, v in cur.items():
                if mode == "reference":
                    ops.accumulate(v, acc[k])                      # train_cnn.py:239 in-place feature sum
                else:
                    ops.accumulate(v.unsqueeze(0), acc[k].unsqueeze(0))
            count += 1
        total = torch.tensor([float(count)], device=next(iter(acc.values())).device)
        dp.allreduce_sums(list(acc.values()) + [total], group)      # C2: one exchange per artist
        length = float(total.item())
        out = {}
        for k, v in acc.items():
            if mode == "reference":
                # gram(sum/len) == gram(sum)/len^2: folds train_cnn.py:242-243's divide into a C x C scale
                g = gram(v.permute(0, 3, 1, 2)) / (length * length)
            else:
                g = v / length
            out[k] = g.expand(batch_size, -1, -1).contiguous()
        return out


class _MSEFunction(torch.autograd.Function):
    """weight * mean((a-b)^2) with the gradient w.r.t. `a` produced in the same pass (nn.MSELoss, train_cnn.py:249,307).

    unit_grad=True promises that the upstream gradient of this loss is exactly 1 (perceptual_step calls
    total.backward() itself with total = content + style): backward then returns the stored gradient as is instead of
    multiplying a feature-map-sized tensor by a device scalar (0.19 ms per step at B=32 for the relu2_2 content term).
    """

    @staticmethod
    def forward(ctx, a, b, weight=1.0, unit_grad=False):
        av = a.detach().permute(0, 2, 3, 1)
        bv = b.detach().permute(0, 2, 3, 1)
        loss = torch.zeros(1, dtype=torch.float32, device=a.device)
        grad = torch.empty(av.shape, dtype=torch.float32, device=a.device) if a.requires_grad else None
        numel = a.numel()
        ops.mse(av, bv, loss, float(weight) / numel, grad, 2.0 * float(weight) / numel)
        ctx.grad, ctx.unit_grad = grad, unit_grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None, None
        ga = ctx.grad if ctx.unit_grad else ctx.grad * g
        return ga.permute(0, 3, 1, 2), None, None, None


class _ContentGramFunction(torch.autograd.Function):
    """relu2_2 feeds BOTH the content loss (train_cnn.py:307) and a style Gram (:321-325).  As two autograd nodes their
    gradients meet in an autograd add over two feature-map-sized tensors (0.12 ms per step at B=32); here one node
    returns (weighted content loss, Gram) and its backward adds the stored content gradient inside the epilogue of the
    Gram-backward convolution (`add` operand of ast_conv_gather).  unit_grad as in _MSEFunction."""

    @staticmethod
    def forward(ctx, f, content_feat, weight, fast, unit_grad):
        b, c, h, w = f.shape
        fv = f.detach()
        if fv.dtype not in (torch.float32, torch.bfloat16):
            fv = fv.float()
        xv = fv.permute(0, 2, 3, 1)
        loss = torch.zeros(1, dtype=torch.float32, device=f.device)
        grad = torch.empty(xv.shape, dtype=torch.float32, device=f.device)
        numel = f.numel()
        ops.mse(xv, content_feat.detach().permute(0, 2, 3, 1), loss, float(weight) / numel, grad, 2.0 * float(weight) / numel)
        g = ops.gram(xv, 1.0 / (c * h * w), tensor=fast and b > 0 and ops.tc_contract_eligible(xv, xv))
        ctx.save_for_backward(fv)
        ctx.grad, ctx.fast, ctx.unit_grad = grad, fast, unit_grad
        return loss[0], g

    @staticmethod
    def backward(ctx, g_loss, dg):
        (fv,) = ctx.saved_tensors
        b, c, h, w = fv.shape
        cgrad = ctx.grad if ctx.unit_grad else ctx.grad * g_loss
        x = fv.permute(0, 2, 3, 1)
        if dg is None:
            return cgrad.permute(0, 3, 1, 2), None, None, None, None
        d = ((dg + dg.transpose(1, 2)) * (1.0 / (c * h * w))).to(fv.dtype).contiguous()
        out = torch.empty((b, h, w, c), dtype=torch.float32, device=fv.device)
        fast = ctx.fast
        ops.conv_gather(x, d.view(b, 1, c, c), cg.conv_fwd(1, 1, 0, h, w), out, add=cgrad, w_img_stride=c * c,
                        tensor=fast and x.is_contiguous() and ops.tc_eligible(x, c), round_tf32=fast)
        return out.permute(0, 3, 1, 2), None, None, None, None


def mse_loss(a, b, weight=1.0):
    """Fused (weighted) MSE forward+gradient for two [B,C,H,W] tensors (b is treated as a constant)."""
    return _MSEFunction.apply(a, b, weight, False)


def perceptual_step(transfer, vgg, content_batch, style_gram, content_weight=CONTENT_WEIGHT,
                    style_weight=STYLE_WEIGHT, backward=True):
    """The loop body of train_cnn.py:295-333 (methods 'random'/'average'/'smartaverage'), without the optimizer.

    Returns (content_loss, style_loss, total_loss) as 0-d device tensors (no host sync).
    Differences from the reference that do not change results: the mean shift is fused into conv1_1's loader,
    and the content branch stops at relu2_2 (the reference computes relu3_3/relu4_3 and discards them).
    """
    shift = neg_mean(content_batch.device)
    generated = transfer(content_batch)                                        # :299
    with torch.no_grad():
        content_feat = vgg(content_batch, shift=shift, upto="relu2_2", only_last=True)["relu2_2"]   # :300
    gen_feats = vgg(generated, shift=shift)                                    # :301
    # :307-308 and the relu2_2 Gram of :321-325 as ONE autograd node (the weight is folded into the kernel; with
    # backward=True the upstream gradient of the content term is exactly 1)
    content_loss, gram22 = _ContentGramFunction.apply(gen_feats["relu2_2"], content_feat, content_weight,
                                                      vgg._mode() == "fast", bool(backward))
    style_loss = 0
    for key, value in gen_feats.items():                                       # :321-325
        g = gram22 if key == "relu2_2" else gram(value, precision=vgg._mode())
        style_loss = style_loss + mse_loss(g.unsqueeze(1), style_gram[key].unsqueeze(1))
    style_loss = style_loss * style_weight
    total = content_loss + style_loss                                          # :329
    if backward:
        total.backward()                                                       # :333
    return content_loss.detach(), style_loss.detach(), total.detach()


class _Prefetched:
    """Handle returned by PerceptualTrainer.prefetch(): a device staging buffer and the event of its H2D copy."""
    __slots__ = ("tensor", "ready", "slot")

    def __init__(self, tensor, ready, slot):
        self.tensor, self.ready, self.slot = tensor, ready, slot


class PerceptualTrainer:
    """Optimizer side of train() (train_cnn.py:247-248,295,334,375) plus data-parallel gradient averaging.

    One process per GPU; when torch.distributed is initialised the TransformerNet gradients are flattened
    into one 1,712,771-float bucket and averaged with a single NCCL all-reduce per step.
    """

    def __init__(self, transfer, vgg, style_gram, lr=LR, weight_decay=1e-4, num_epochs=200, num_steps=2,
                 content_weight=CONTENT_WEIGHT, style_weight=STYLE_WEIGHT, group=None, cuda_graph=False):
        self.transfer, self.vgg, self.style_gram = transfer, vgg, style_gram
        self.content_weight, self.style_weight = content_weight, style_weight
        self.params = [p for p in transfer.parameters()]
        on_cuda = self.params[0].is_cuda
        # fused=True: torch's single-kernel Adam (same update rule, L2 weight decay) instead of ~15 foreach kernels per step
        self.optimizer = torch.optim.Adam(self.params, lr=lr, weight_decay=weight_decay, fused=bool(on_cuda),
                                          capturable=bool(cuda_graph and on_cuda))                 # :247
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=max(1, num_epochs // num_steps),
                                                         gamma=0.5)                                # :248
        self.group = group
        self._flat = None
        # cuda_graph=True: after 3 eager warm-up steps the whole step (fwd, bwd, all-reduce, Adam: ~350 launches) is
        # captured once per input shape and replayed, removing host launch overhead and inter-kernel gaps.
        self.cuda_graph = bool(cuda_graph and on_cuda)
        self._graph, self._static_in, self._static_losses, self._eager_steps = None, None, None, 0
        self._copy_stream = None

    # ---- host -> device input pipeline --------------------------------------------------------------------------
    def prefetch(self, host_batch):
        """Start copying a (pinned) host batch to the device on a side stream and return a handle that `step()` accepts.

        Called for batch i+1 before `step(batch i)` is waited on, the PCIe copy (25 MB at B=32, 256^2) overlaps the
        compute of step i instead of preceding step i+1.  Two device staging buffers alternate; a buffer is only
        overwritten after the step that consumed it has copied it into the step's input.
        """
        dev = self.params[0].device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage, self._stage_free, self._stage_i = [None, None], [None, None], 0
        i = self._stage_i
        self._stage_i ^= 1
        if self._stage[i] is None or self._stage[i].shape != host_batch.shape:
            self._stage[i] = torch.empty(host_batch.shape, dtype=torch.float32, device=dev)
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[i] is not None:
                self._copy_stream.wait_event(self._stage_free[i])
            self._stage[i].copy_(host_batch, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
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
        if self._flat is None:
            self._flat = dp.GradBucket()
        self._flat.allreduce_mean(self.params, self.group)

    def _eager_step(self, content_batch):
        self.optimizer.zero_grad(set_to_none=True)                                                 # :295
        losses = perceptual_step(self.transfer, self.vgg, content_batch, self.style_gram,
                                 self.content_weight, self.style_weight, backward=True)
        self._allreduce_grads()
        self.optimizer.step()                                                                      # :334
        return losses

    def _capture(self, content_batch):
        self._static_in = content_batch.clone()
        self._graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self._graph):
            self._static_losses = self._eager_step(self._static_in)

    def step(self, content_batch):
        """One optimisation step on `content_batch`: a device tensor, a host tensor, or a handle from `prefetch()`."""
        content_batch, slot = self._consume(content_batch)
        if not self.cuda_graph:
            if not content_batch.is_cuda:
                content_batch = content_batch.to(self.params[0].device, non_blocking=True)
            losses = self._eager_step(content_batch)
            self._release(slot)
            return losses
        if self._graph is not None and self._static_in.shape == content_batch.shape:
            self._static_in.copy_(content_batch, non_blocking=True)
            self._release(slot)
            self._graph.replay()
            return self._static_losses
        if not content_batch.is_cuda:
            content_batch = content_batch.to(self.params[0].device, non_blocking=True)
        if self._eager_steps < 3:                       # warm up caches (packed weights, tap tables, allocator)
            self._eager_steps += 1
            losses = self._eager_step(content_batch)
            self._release(slot)
            return losses
        self._capture(content_batch)                    # capture does not execute: run the graph once for this batch
        self._release(slot)
        self._graph.replay()
        return self._static_losses

    def end_epoch(self):
        self.scheduler.step()                                                                      # :375
