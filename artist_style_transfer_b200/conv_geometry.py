"""Tap tables that turn every convolution-like op of the path into `ast_conv_gather` launches.

Pure Python (no torch, no CUDA) so the geometry is unit-tested on CPU against torch.nn.functional.

One `Launch` = one call of the gather primitive (include/ast.h):
    out[n, oy0 + so*i, ox0 + so*j, co] = sum_t sum_ci in[n, si*i + dy[t], si*j + dx[t], ci] * Wt[t][co][ci]
`wtaps[t] = (u, v)` is the kernel position whose weights multiply tap t.

Ops (reference call sites): Conv2d after ReflectionPad2d (cnn.py:58,63), its data gradient, ConvTranspose2d
k3 s2 p1 op1 / k1 s1 (cnn.py:107-109) and its data gradient, VGG Conv2d 3x3 pad 1 (train_cnn.py:54).
"""
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass
class Launch:
    mi: int
    mj: int
    si: int
    so: int
    oy0: int
    ox0: int
    taps: List[Tuple[int, int]] = field(default_factory=list)    # (dy, dx)
    wtaps: List[Tuple[int, int]] = field(default_factory=list)   # (u, v) kernel position per tap
    woff: int = 0                                                # first tap of this launch in the packed weights


def conv_out_size(n_in, k, stride, pad):
    return (n_in + 2 * pad - k) // stride + 1


def convT_out_size(n_in, k, stride, pad, output_padding):
    return (n_in - 1) * stride - 2 * pad + k + output_padding


def conv_fwd(k, stride, pad, h_in, w_in):
    """y[i,j] = sum_{u,v} x[s*i+u-pad, s*j+v-pad] W[u,v]  (cross-correlation, like nn.Conv2d)."""
    ho, wo = conv_out_size(h_in, k, stride, pad), conv_out_size(w_in, k, stride, pad)
    taps = [(u - pad, v - pad) for u in range(k) for v in range(k)]
    wtaps = [(u, v) for u in range(k) for v in range(k)]
    return [Launch(ho, wo, stride, 1, 0, 0, taps, wtaps, 0)]


def conv_dgrad(k, stride, pad, h_in, w_in):
    """gx[y,x] = sum over (i,u): s*i+u-pad == y of gy[i,j] W[u,v]; one launch per output phase (y mod s, x mod s).

    Also the forward of ConvTranspose2d(k, stride, pad) producing an (h_in, w_in) image.
    """
    launches, woff = [], 0
    for py in range(stride):
        for px in range(stride):
            us = [u for u in range(k) if (py + pad - u) % stride == 0]
            vs = [v for v in range(k) if (px + pad - v) % stride == 0]
            mi = (h_in - py + stride - 1) // stride
            mj = (w_in - px + stride - 1) // stride
            if mi <= 0 or mj <= 0:
                continue
            if not us or not vs:
                raise ValueError("phase without taps (k < stride) is not supported")
            taps = [((py + pad - u) // stride, (px + pad - v) // stride) for u in us for v in vs]
            wtaps = [(u, v) for u in us for v in vs]
            launches.append(Launch(mi, mj, 1, stride, py, px, taps, wtaps, woff))
            woff += len(taps)
    return launches


def convT_fwd(k, stride, pad, output_padding, h_in, w_in):
    ho = convT_out_size(h_in, k, stride, pad, output_padding)
    wo = convT_out_size(w_in, k, stride, pad, output_padding)
    return conv_dgrad(k, stride, pad, ho, wo)


def convT_dgrad(k, stride, pad, h_out, w_out):
    """gx[i,j] = sum_{u,v} gy[s*i+u-pad, s*j+v-pad] W[u,v]: same geometry as conv_fwd on the gradient image."""
    return conv_fwd(k, stride, pad, h_out, w_out)


def all_wtaps(launches):
    out = []
    for l in launches:
        out.extend(l.wtaps)
    return out
