"""B200-native mirror of the reference transform network (`/root/reference/cnn.py`).

Same class names, constructor arguments, forward signatures and `state_dict` keys as the reference
(`StyleTransfer` cnn.py:10-49, `ConvLayer` :52-79, `ResidualLayer` :82-99, `DeconvLayer` :102-124), so
reference checkpoints load with `strict=True`.  Underneath, every layer is a sequence of calls into
libast_b200.so (hand-written sm_100a kernels); there is no PyTorch-op or CPU fallback.

Internal data flow: activations are NHWC; each stage is
    gather-conv (raw, pre-norm)  ->  InstanceNorm stats  ->  fused apply (+ReLU, +residual) that writes the
    next conv's reflection-padded input buffer directly.
The backward pass runs the same stages in reverse inside one autograd.Function.
"""
import math
import os

import torch
import torch.nn as nn

from . import conv_geometry as cg
from . import ops

# precision modes (north_star): 'fp32' = strict, FFMA kernels, fp32 activations;
# 'fast' = bf16 activations/weights for the transform net on tcgen05 where the shape is supported.
_MODES = ("fp32", "fast")
_default_mode = "fp32"


def set_default_precision(mode):
    global _default_mode
    if mode not in _MODES:
        raise ValueError(f"precision must be one of {_MODES}")
    _default_mode = mode


def get_default_precision():
    return _default_mode


def _thin_wgrad_mode():
    """'taps' (default): 9-tap tap-group launch of the contraction kernel (measured 0.33 ms/step faster at B=32);
    'unfold': vertical taps folded into channels by ast_unfold_rows, then a single-tap contraction."""
    import os
    return os.environ.get("AST_THIN_WGRAD", "taps")


def _fused_stats():
    import os
    return os.environ.get("AST_FUSED_STATS", "1") == "1"


def _grad_fp32():
    import os
    return os.environ.get("AST_GRAD_FP32", "0") == "1"


class _ConvParams(nn.Module):
    """Parameter holder laid out like nn.Conv2d / nn.ConvTranspose2d (weight, bias) with PyTorch's default init."""

    def __init__(self, weight_shape, fan_in, bias_size):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(weight_shape))
        self.bias = nn.Parameter(torch.empty(bias_size))
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)   # == kaiming_uniform_(a=sqrt(5))
            self.bias.uniform_(-bound, bound)


class _NormParams(nn.Module):
    """Parameter holder of nn.InstanceNorm2d(affine=True): weight=1, bias=0, no running stats."""

    def __init__(self, channels):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))


class _Stage:
    __slots__ = ("kind", "cin", "cout", "k", "stride", "opad", "norm", "relu", "res_from", "conv", "normp")

    def __init__(self, kind, cin, cout, k, stride, opad, norm, relu, res_from, conv, normp):
        self.kind, self.cin, self.cout, self.k, self.stride, self.opad = kind, cin, cout, k, stride, opad
        self.norm, self.relu, self.res_from, self.conv, self.normp = norm, relu, res_from, conv, normp

    @property
    def in_pad(self):
        return self.k // 2 if (self.kind == "conv" and self.k > 1) else 0


def _interior(t, p):
    return t if p == 0 else t[:, p:t.shape[1] - p, p:t.shape[2] - p, :]


def _fwd_geometry(st, hin, win):
    if st.kind == "conv":
        ls = cg.conv_fwd(st.k, st.stride, 0, hin, win)
        return ls, ls[0].mi, ls[0].mj
    ho = cg.convT_out_size(hin, st.k, st.stride, st.k // 2, st.opad)
    wo = cg.convT_out_size(win, st.k, st.stride, st.k // 2, st.opad)
    return cg.convT_fwd(st.k, st.stride, st.k // 2, st.opad, hin, win), ho, wo


def _pack_fwd(st, w, launches, dtype):
    k2 = st.k * st.k
    if st.kind == "conv":      # (Co,Ci,k,k) -> [t][co][ci]
        return ops.pack_weights(w, launches, st.cout, st.cin, st.cin * k2, k2, st.k, 1, dtype)
    return ops.pack_weights(w, launches, st.cout, st.cin, k2, st.cout * k2, st.k, 1, dtype)  # (Ci,Co,k,k)


def _pack_dgrad(st, w, launches, dtype):
    k2 = st.k * st.k
    if st.kind == "conv":      # [t][ci][co]
        return ops.pack_weights(w, launches, st.cin, st.cout, k2, st.cin * k2, st.k, 1, dtype)
    return ops.pack_weights(w, launches, st.cin, st.cout, st.cout * k2, k2, st.k, 1, dtype)


# ---- 3-channel ends on the tensor cores (fast mode): the kw horizontal taps of the 9x9 filters are folded into
# channels by ops.row_im2col so that TMA/tcgen05 see 32-channel NHWC tensors (csrc/im2col.cu explains the layouts).
def _thin_in(st):
    return st.kind == "conv" and st.stride == 1 and st.k > 1 and st.cin * st.k <= 32 and st.cout % 32 == 0 and st.norm


def _thin_out(st):
    return st.kind == "conv" and st.stride == 1 and st.k > 1 and st.cout * st.k <= 32 and st.cin % 32 == 0 and not st.norm


def _vtaps(k, sign=1):
    taps = [(sign * dy, 0) for dy in range(k)]
    return taps, [(dy, 0) for dy in range(k)]


def _thin_in_launch(st, h, w):
    taps, wt = _vtaps(st.k)
    return [cg.Launch(h, w, 1, 1, 0, 0, taps, wt, 0)]


def _thin_in_pack(st, w, dtype):        # [dy][co][dx*cin + c] = W[co][c][dy][dx]
    k, k2 = st.k, st.k * st.k
    return ops.pack_weights_ex(w, [dy * k for dy in range(k)], st.cout, st.cout, 32, k * st.cin, st.cin,
                               st.cin * k2, 1, k2, dtype)


def _thin_out_pack_fwd(st, w, launches, dtype):   # [t][co (3 of 32)][c] = W[co][c][u][v]
    k2 = st.k * st.k
    offs = [u * st.k + v for u, v in cg.all_wtaps(launches)]
    return ops.pack_weights_ex(w, offs, 32, st.cout, st.cin, st.cin, 1, st.cin * k2, k2, 0, dtype)


def _thin_out_pack_vfwd(st, w, dtype):   # [dy][dx*cout + co (k*cout of 32)][c] = W[co][c][dy][dx]
    k = st.k
    w2 = w.permute(2, 3, 0, 1).contiguous()            # (dy, dx, co, c): the row index dx*cout+co becomes linear
    return ops.pack_weights_ex(w2, [dy * k * st.cout * st.cin for dy in range(k)], 32, k * st.cout, st.cin, st.cin,
                               st.cin, st.cin, 0, 1, dtype)


def _thin_out_fold():
    """AST_THIN_OUT=taps: the last 9x9 layer as one 81-tap conv; default: 9 vertical taps + ast_fold_rows."""
    return os.environ.get("AST_THIN_OUT", "fold") != "taps"


def _thin_out_pack_dgrad(st, w, dtype):   # [dy][c][dx*cout + co] = W[co][c][dy][dx]
    k, k2 = st.k, st.k * st.k
    return ops.pack_weights_ex(w, [dy * k for dy in range(k)], st.cin, st.cin, 32, k * st.cout, st.cout,
                               k2, 1, st.cin * k2, dtype)


class _ZeroArena:
    """One zero-filled fp32 buffer per forward / backward pass, carved into the many small accumulators (InstanceNorm
    sums, filter-gradient scratch, dead conv-bias gradients): one fill kernel instead of ~50 per step."""

    def __init__(self, sizes, device):
        self.buf = torch.zeros(sum((s + 3) // 4 * 4 for s in sizes), dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape):
        numel = 1
        for d in shape:
            numel *= d
        view = self.buf[self.off:self.off + numel].view(*shape)
        self.off += (numel + 3) // 4 * 4          # keep every slice 16-byte aligned (vector atomics)
        assert self.off <= self.buf.numel()
        return view


class _StageFunction(torch.autograd.Function):
    """Forward/backward of a list of stages as ONE autograd node (x: NCHW fp32 in, NCHW fp32 out)."""

    @staticmethod
    def forward(ctx, x, stages, mode, *params):
        if not x.is_cuda:
            raise RuntimeError("StyleTransfer kernels run on CUDA only (no CPU fallback); move the module and "
                               "input to a B200")
        adt = torch.bfloat16 if mode == "fast" else torch.float32
        x = x.detach().to(torch.float32)
        n, _, h, w = x.shape
        dev = x.device
        pit = iter(params)
        P = []
        for st in stages:
            cw, cb = next(pit), next(pit)
            P.append((cw, cb, next(pit), next(pit)) if st.norm else (cw, cb, None, None))
        p0 = stages[0].in_pad
        thin_in = mode == "fast" and _thin_in(stages[0]) and ops.tc_eligible(torch.empty(0, 1, 1, 32, dtype=adt, device=dev), stages[0].cout)
        if thin_in:     # row-im2col of the reflect-padded image: [N, H+2p, W, 32] (k*cin channels used)
            node = torch.empty((n, h + 2 * p0, w, 32), dtype=adt, device=dev)
            ops.row_im2col(x.permute(0, 2, 3, 1), node, stages[0].k, 1, p0, p0, True)
        else:
            node = torch.empty((n, h + 2 * p0, w + 2 * p0, stages[0].cin), dtype=adt, device=dev)
            ops.copy_image(x.permute(0, 2, 3, 1), node, pad=p0)
        nodes, node_pad, saved = [node], [p0], []
        out = None
        zeros = _ZeroArena([n * st.cout * 2 for st in stages if st.norm], dev)
        for i, st in enumerate(stages):
            cw, cb, gam, bet = P[i]
            xin = nodes[-1]
            last = i == len(stages) - 1
            thin_out = mode == "fast" and _thin_out(st) and ops.tc_eligible(xin, 32)
            if i == 0 and thin_in:
                launches, ho, wo = _thin_in_launch(st, h, w), h, w
                wp = _thin_in_pack(st, cw.detach(), adt)
            else:
                launches, ho, wo = _fwd_geometry(st, xin.shape[1], xin.shape[2])
                if thin_out and _thin_out_fold():
                    wp = None                      # packed for the vertical-tap formulation below
                else:
                    wp = (_thin_out_pack_fwd(st, cw.detach(), launches, adt) if thin_out
                          else _pack_fwd(st, cw.detach(), launches, adt))
            if not st.norm:
                assert last, "a stage without norm must be the last one"
                out = torch.empty((n, st.cout, ho, wo), dtype=torch.float32, device=dev)
                if thin_out and _thin_out_fold():
                    # k vertical taps on the tensor cores give, per pixel, the partial sums of all k horizontal taps
                    # as k*cout (27 of 32) channels; ast_fold_rows adds the k shifted partials and the bias
                    taps, wt = _vtaps(st.k)
                    part = torch.empty((n, ho, xin.shape[2], 32), dtype=torch.float32, device=dev)
                    ops.conv_gather(xin, _thin_out_pack_vfwd(st, cw.detach(), adt),
                                    [cg.Launch(ho, xin.shape[2], 1, 1, 0, 0, taps, wt, 0)], part, tensor=True)
                    ops.fold_rows(part, out.permute(0, 2, 3, 1), st.k, bias=cb.detach(), relu=st.relu)
                    del part
                else:
                    ops.conv_gather(xin, wp, launches, out.permute(0, 2, 3, 1), bias=cb.detach(), relu=st.relu,
                                    tensor=thin_out)
                saved.append((launches, None, None, None))
            else:
                raw = torch.empty((n, ho, wo, st.cout), dtype=adt, device=dev)
                # conv bias is dead under InstanceNorm and is not added
                use_tc = mode == "fast" and ops.tc_eligible(xin, st.cout)
                if use_tc and _fused_stats():      # sum x / sum x^2 accumulated by the conv epilogue: no stats pass
                    sums = zeros.take(n * st.cout * 2)
                    ops.conv_gather(xin, wp, launches, raw, tensor=True, stats=sums)
                    mean, rstd = ops.instnorm_finalize(sums, n, st.cout, ho * wo)
                else:
                    ops.conv_gather(xin, wp, launches, raw, tensor=use_tc)
                    mean, rstd = ops.instnorm_stats(raw)
                pn = 0 if last else stages[i + 1].in_pad
                post = torch.empty((n, ho + 2 * pn, wo + 2 * pn, st.cout), dtype=adt, device=dev)
                res = None
                if st.res_from is not None:
                    j = st.res_from + 1
                    res = _interior(nodes[j], node_pad[j])
                ops.instnorm_apply(raw, mean, rstd, gam.detach(), bet.detach(), post, pn, st.relu, residual=res)
                nodes.append(post)
                node_pad.append(pn)
                saved.append((launches, raw, mean, rstd))
                if last:
                    out = torch.empty((n, st.cout, ho, wo), dtype=torch.float32, device=dev)
                    ops.copy_image(post, out.permute(0, 2, 3, 1))
        ctx.stages, ctx.mode, ctx.P = stages, mode, P
        ctx.nodes, ctx.node_pad, ctx.saved = nodes, node_pad, saved
        ctx.thin_in = thin_in
        return out

    @staticmethod
    def backward(ctx, gout):
        stages, P, nodes, node_pad, saved = ctx.stages, ctx.P, ctx.nodes, ctx.node_pad, ctx.saved
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("gradient w.r.t. the input image is not part of the training path "
                                      "(train_cnn.py:298-299 feeds data that does not require grad)")
        adt = nodes[0].dtype
        # dtype of the gradients entering an InstanceNorm backward (dgrad outputs, residual skip gradients).  The
        # g - mean(g) - xhat*mean(g*xhat) cancellation amplifies their rounding error, so they can be kept in fp32
        # (AST_GRAD_FP32=1) at the cost of 2x traffic on those tensors; see DESIGN.md "gradient noise".
        gdt = torch.float32 if (adt == torch.float32 or _grad_fp32()) else adt
        gout = gout.to(torch.float32)
        L = len(stages)
        gpad = [None] * (L + 1)
        gextra = [None] * (L + 1)
        grads = []
        zeros = _ZeroArena([max(st.k * st.k * st.cout * st.cin, st.k * 32 * max(st.cout, st.cin)) + st.cout
                            for st in stages], gout.device)
        # per-(n,c) InstanceNorm-backward sums of all layers with the same channel count share one buffer
        # [layer][2][n][c], so dbeta / dgamma of those layers come out of ONE batch reduction at the end
        nb = gout.shape[0]
        bank_rows = {}
        for st in stages:
            if st.norm:
                bank_rows[st.cout] = bank_rows.get(st.cout, 0) + 1
        banks = {c: torch.empty((rows, 2, nb, c), dtype=torch.float32, device=gout.device) for c, rows in bank_rows.items()}
        bank_next = {c: 0 for c in bank_rows}
        bank_slot = {}
        for i in reversed(range(L)):
            st = stages[i]
            cw, cb, gam, bet = P[i]
            launches, raw, mean, rstd = saved[i]
            xin = nodes[i]
            thin_out = ctx.mode == "fast" and _thin_out(st) and ops.tc_eligible(xin, 32)
            if not st.norm:
                d_raw = gout.permute(0, 2, 3, 1)
                g_cb = gout.sum(dim=(0, 2, 3))
                g_gam = g_bet = None
                if thin_out:   # Dr[n,y,x',dx*cout+co] = dOut[n,y,x'-dx,co]: shared by the wgrad and the dgrad
                    pk = st.k // 2
                    nb, hb, wb = gout.shape[0], gout.shape[2], gout.shape[3]
                    d_raw = torch.empty((nb, hb, wb + 2 * pk, 32), dtype=adt, device=gout.device)
                    ops.row_im2col(gout.permute(0, 2, 3, 1), d_raw, st.k, -1, 0, 0, False)
            else:
                j = i + 1
                n, ho, wo, c = raw.shape
                if i == L - 1:
                    ge = torch.empty((n, ho, wo, c), dtype=torch.float32, device=raw.device)
                    ops.copy_image(gout.permute(0, 2, 3, 1), ge)
                    gextra[j] = ge
                d_raw = torch.empty_like(raw)
                gtotal = torch.empty(raw.shape, dtype=gdt, device=raw.device) if st.res_from is not None else None
                slot = bank_next[c]
                bank_next[c] += 1
                bank_slot[i] = (c, slot)
                ops.instnorm_bwd(raw, mean, rstd, gam.detach(), bet.detach(), gpad[j], node_pad[j],
                                 gextra[j], st.relu, d_raw, gtotal, s12=banks[c][slot].view(2, n * c))
                g_bet = g_gam = None                 # filled from the bank reductions after the loop
                g_cb = zeros.take(st.cout)           # exactly zero under InstanceNorm
                if gtotal is not None:
                    assert gextra[st.res_from + 1] is None
                    gextra[st.res_from + 1] = gtotal
                gpad[j] = gextra[j] = None           # free
            k2 = st.k * st.k
            wtc = ctx.mode == "fast" and ops.tc_contract_eligible(xin, d_raw)
            one_tap = [cg.Launch(1, 1, 1, 1, 0, 0, [(0, 0)], [(0, 0)], 0)]
            thin_taps = _thin_wgrad_mode() == "taps"
            if i == 0 and ctx.thin_in and thin_taps:
                # k vertical taps handled as a tap group inside the contraction kernel (d_raw loaded once per stage):
                # tmp[dy][co][dx*cin+c] += sum_p dY[p][co] * Xr[p + dy rows][dx*cin+c]
                tmp = zeros.take(st.k, st.cout, 32)
                ops.wgrad_gather(xin, d_raw, launches, tmp, 32, 1, st.cout * 32, 0, tensor=wtc)
                g_cw = tmp[:, :, :st.k * st.cin].reshape(st.k, st.cout, st.k, st.cin).permute(1, 3, 0, 2).contiguous()
            elif not st.norm and thin_out and thin_taps:
                # tmp[dy][dx*cout+co][c] += sum_{y,x'} Dr[y][x'][dx*cout+co] * xin[y+dy][x'][c]
                taps, wt = _vtaps(st.k)
                lw = [cg.Launch(d_raw.shape[1], d_raw.shape[2], 1, 1, 0, 0, taps, wt, 0)]
                tmp = zeros.take(st.k, 32, st.cin)
                ops.wgrad_gather(xin, d_raw, lw, tmp, st.cin, 1, 32 * st.cin, 0, tensor=True)
                g_cw = tmp[:, :st.k * st.cout, :].reshape(st.k, st.k, st.cout, st.cin).permute(2, 3, 0, 1).contiguous()
            elif i == 0 and ctx.thin_in:
                # fold the k vertical taps into channels too (X[y][x][dy*32 + dx*cin+c] = Xr[y+dy][x][dx*cin+c], 16-byte
                # copies), then ONE single-tap contraction: tmp[co][dy*32 + dx*cin+c] = sum_p dY[p][co] * X[p][...]
                n_, h_, w_ = d_raw.shape[0], d_raw.shape[1], d_raw.shape[2]
                xf = torch.empty((n_, h_, w_, st.k * 32), dtype=adt, device=xin.device)
                ops.unfold_rows(xin, xf, st.k, 1)
                one_tap[0].mi, one_tap[0].mj = h_, w_
                # operands swapped (the 288-channel tensor provides the M rows): tmp[dy*32 + dx*cin+c][co]
                tmp = zeros.take(st.k * 32, st.cout)
                ops.wgrad_gather(d_raw, xf, one_tap, tmp, st.cout, 1, 0, 0, tensor=wtc)
                g_cw = (tmp.view(st.k, 32, st.cout)[:, :st.k * st.cin, :].reshape(st.k, st.k, st.cin, st.cout)
                        .permute(3, 2, 0, 1).contiguous())
                del xf
            elif not st.norm and thin_out:
                # D[y'][x'][dy*32 + dx*cout+co] = Dr[y'-dy][x'][dx*cout+co] (zero outside), then ONE single-tap
                # contraction against the padded input: tmp[dy*32 + dx*cout+co][c] = sum_{y',x'} D[..] * xin[y'][x'][c]
                df = torch.empty((xin.shape[0], xin.shape[1], xin.shape[2], st.k * 32), dtype=adt, device=xin.device)
                ops.unfold_rows(d_raw, df, st.k, -1)
                one_tap[0].mi, one_tap[0].mj = xin.shape[1], xin.shape[2]
                tmp = zeros.take(st.k * 32, st.cin)
                ops.wgrad_gather(xin, df, one_tap, tmp, st.cin, 1, 0, 0, tensor=True)
                g_cw = (tmp.view(st.k, 32, st.cin)[:, :st.k * st.cout, :].reshape(st.k, st.k, st.cout, st.cin)
                        .permute(2, 3, 0, 1).contiguous())
                del df
            elif wtc:
                # tap-major scratch tmp[tap][co][ci]: ci contiguous, so the contraction epilogue reduces with 16-byte
                # vector atomics; the permute to the parameter layout is one small copy
                tmp = zeros.take(k2, st.cout, st.cin)
                ops.wgrad_gather(xin, d_raw, launches, tmp, st.cin, 1, st.k * st.cout * st.cin, st.cout * st.cin,
                                 tensor=True)
                tmp = tmp.view(st.k, st.k, st.cout, st.cin)
                g_cw = (tmp.permute(2, 3, 0, 1) if st.kind == "conv" else tmp.permute(3, 2, 0, 1)).contiguous()
            else:
                g_cw = torch.zeros_like(cw, dtype=torch.float32)
                if st.kind == "conv":
                    ops.wgrad_gather(xin, d_raw, launches, g_cw, st.cin * k2, k2, st.k, 1, tensor=wtc)
                else:
                    ops.wgrad_gather(xin, d_raw, launches, g_cw, k2, st.cout * k2, st.k, 1, tensor=wtc)
            if i > 0 and not st.norm and thin_out:
                taps, wt = _vtaps(st.k, -1)
                dl = [cg.Launch(xin.shape[1], xin.shape[2], 1, 1, 0, 0, taps, wt, 0)]
                g_in = torch.empty(xin.shape, dtype=gdt, device=xin.device)
                ops.conv_gather(d_raw, _thin_out_pack_dgrad(st, cw.detach(), adt), dl, g_in, tensor=True)
                gpad[i] = g_in
            elif i > 0:
                if st.kind == "conv":
                    dl = cg.conv_dgrad(st.k, st.stride, 0, xin.shape[1], xin.shape[2])
                else:
                    dl = cg.convT_dgrad(st.k, st.stride, st.k // 2, d_raw.shape[1], d_raw.shape[2])
                wpd = _pack_dgrad(st, cw.detach(), dl, adt)
                g_in = torch.empty(xin.shape, dtype=gdt, device=xin.device)
                src = d_raw
                if src.dtype != adt:               # fp32 NCHW grad of the last conv feeding a bf16 dgrad
                    src = torch.empty(d_raw.shape, dtype=adt, device=xin.device)
                    ops.copy_image(d_raw, src)
                ops.conv_gather(src, wpd, dl, g_in, tensor=ctx.mode == "fast" and ops.tc_eligible(src, st.cin))
                gpad[i] = g_in
            grads.append([g_cw, g_cb, g_gam, g_bet, i])
        reduced = {c: bank.sum(2) for c, bank in banks.items()}       # [layer][2][c]: dbeta = [.,0], dgamma = [.,1]
        for gr in grads:
            if gr[4] in bank_slot:
                c, slot = bank_slot[gr[4]]
                gr[3], gr[2] = reduced[c][slot, 0], reduced[c][slot, 1]
        grads = [tuple(gr[:4]) for gr in grads]
        grads.reverse()
        flat = []
        for st, (g_cw, g_cb, g_gam, g_bet) in zip(stages, grads):
            flat += [g_cw, g_cb]
            if st.norm:
                flat += [g_gam, g_bet]
        ctx.nodes = ctx.saved = None
        return (None, None, None, *flat)


def _run_stages(x, stages, mode):
    params = []
    for st in stages:
        params += [st.conv.weight, st.conv.bias]
        if st.norm:
            params += [st.normp.weight, st.normp.bias]
    return _StageFunction.apply(x, tuple(stages), mode, *params)


class _Precision:
    precision = None  # None -> module default (set_default_precision)

    def _mode(self):
        return self.precision or _default_mode


class ConvLayer(nn.Module, _Precision):
    """ReflectionPad(k//2) -> Conv2d(k, stride) -> InstanceNorm2d(affine) (cnn.py:52-79)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, norm="instance"):
        super().__init__()
        if norm not in ("instance", "None"):
            raise ValueError("only norm='instance' or 'None' is on the accelerated path (cnn.py:66-70 'batch' is unused)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.norm_type = kernel_size, stride, norm
        self.conv_layer = _ConvParams((out_channels, in_channels, kernel_size, kernel_size),
                                      in_channels * kernel_size * kernel_size, out_channels)
        if norm == "instance":
            self.norm_layer = _NormParams(out_channels)

    def _stage(self, relu=False, res_from=None):
        return _Stage("conv", self.in_channels, self.out_channels, self.kernel_size, self.stride, 0,
                      self.norm_type == "instance", relu, res_from, self.conv_layer,
                      getattr(self, "norm_layer", None))

    def forward(self, x):
        return _run_stages(x, [self._stage()], self._mode())


class ResidualLayer(nn.Module, _Precision):
    """conv2(relu(conv1(x))) + x, no activation after the add (cnn.py:82-99)."""

    def __init__(self, channels=128, kernel_size=3):
        super().__init__()
        self.conv1 = ConvLayer(channels, channels, kernel_size, stride=1)
        self.relu = nn.ReLU()
        self.conv2 = ConvLayer(channels, channels, kernel_size, stride=1)

    def _stages(self, base):
        """`base` = index of the stage whose output is this block's input (-1 = the run's input)."""
        return [self.conv1._stage(relu=True), self.conv2._stage(relu=False, res_from=base)]

    def forward(self, x):
        return _run_stages(x, self._stages(-1), self._mode())


class DeconvLayer(nn.Module, _Precision):
    """ConvTranspose2d(k, stride, padding=k//2, output_padding) -> InstanceNorm2d(affine) (cnn.py:102-124)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, output_padding, norm="instance"):
        super().__init__()
        if norm != "instance":
            raise ValueError("only norm='instance' is on the accelerated path")
        if (kernel_size, stride, output_padding) not in ((1, 1, 0), (3, 2, 1)):
            raise ValueError("DeconvLayer supports (k,stride,output_padding) = (1,1,0) or (3,2,1) (cnn.py:33-37)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.output_padding, self.norm_type = kernel_size, stride, output_padding, norm
        self.conv_transpose = _ConvParams((in_channels, out_channels, kernel_size, kernel_size),
                                          out_channels * kernel_size * kernel_size, out_channels)
        self.norm_layer = _NormParams(out_channels)

    def _stage(self, relu=False):
        return _Stage("deconv", self.in_channels, self.out_channels, self.kernel_size, self.stride,
                      self.output_padding, True, relu, None, self.conv_transpose, self.norm_layer)

    def forward(self, x):
        return _run_stages(x, [self._stage()], self._mode())


class StyleTransfer(nn.Module, _Precision):
    """Johnson-style image transform net with the reference's two extra 1x1 layers (cnn.py:10-49).

    `StyleTransfer(state_dict_filename=None, device=None, precision=None)`; master parameters stay fp32
    (the reference's `.double()` at cnn.py:43 is not reproduced: fp64 has no tensor-core path).
    """

    def __init__(self, state_dict_filename=None, device=None, precision=None):
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        super().__init__()
        self.precision = precision
        self.ConvBlock = nn.Sequential(
            ConvLayer(3, 32, 9, 1), nn.ReLU(),
            ConvLayer(32, 64, 3, 2), nn.ReLU(),
            ConvLayer(64, 128, 3, 2), nn.ReLU(),
            ConvLayer(128, 128, 1, 1), nn.ReLU())
        self.ResidualBlock = nn.Sequential(*[ResidualLayer(128, 3) for _ in range(5)])
        self.DeconvBlock = nn.Sequential(
            DeconvLayer(128, 128, 1, 1, 0), nn.ReLU(),
            DeconvLayer(128, 64, 3, 2, 1), nn.ReLU(),
            DeconvLayer(64, 32, 3, 2, 1), nn.ReLU(),
            ConvLayer(32, 3, 9, 1, norm="None"))
        if state_dict_filename is not None:
            self.load_state_dict(torch.load(state_dict_filename, map_location=device), strict=True)
        self.to(device)

    def _stages(self):
        stages = []
        for block in (self.ConvBlock, self.ResidualBlock, self.DeconvBlock):
            mods = list(block)
            for idx, m in enumerate(mods):
                nxt_relu = idx + 1 < len(mods) and isinstance(mods[idx + 1], nn.ReLU)
                if isinstance(m, (ConvLayer, DeconvLayer)):
                    stages.append(m._stage(relu=nxt_relu))
                elif isinstance(m, ResidualLayer):
                    stages += m._stages(len(stages) - 1)
        return stages

    def forward(self, x):
        return _run_stages(x, self._stages(), self._mode())


TransformerNet = StyleTransfer  # north_star's name for the same network
