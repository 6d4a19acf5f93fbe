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


# ---------------------------------------------------------------------------------------------------------------------
# Block-stacked launches (ast_conv_stacked, include/ast.h): nblk = 128 / cout blocks of outputs share every MMA.

@dataclass
class Stacked:
    nblk: int
    mi: int
    mj: int
    sy: int                        # input rows per grid row
    soy: int                       # output steps per grid row / column
    sox: int
    oy: List[int] = field(default_factory=list)                   # output origin per block
    ox: List[int] = field(default_factory=list)
    vt: List[Tuple[int, int]] = field(default_factory=list)       # (dy, dx) input shift of each virtual tap
    src: List[List[object]] = field(default_factory=list)         # src[v][g] = kernel position (u, v) or None (zero rows)

    @property
    def ntaps(self):
        return sum(1 for row in self.src for s in row if s is not None)

    def places(self):
        """{kernel position: [(virtual tap, block), ...]}: where each filter tap sits in the stacked filter."""
        out = {}
        for v, row in enumerate(self.src):
            for g, s in enumerate(row):
                if s is not None:
                    out.setdefault(s, []).append((v, g))
        return out


def stack_phases(launches):
    """The sub-pixel phases of a stride-2 ConvTranspose2d forward / stride-2 conv data gradient (all with si = 1 and the
    same so) as ONE stacked launch: block g = phase g, virtual taps = the union of the phases' input shifts."""
    assert len(launches) in (2, 4) and all(l.si == 1 and l.so == launches[0].so for l in launches)
    vt = sorted({t for l in launches for t in l.taps})
    src = [[dict(zip(l.taps, l.wtaps)).get(t) for l in launches] for t in vt]
    return Stacked(len(launches), max(l.mi for l in launches), max(l.mj for l in launches), 1, launches[0].so,
                   launches[0].so, [l.oy0 for l in launches], [l.ox0 for l in launches], vt, src)


def stack_rows(launch, nblk):
    """A stride-1 launch with its output rows interleaved over nblk blocks: block g owns rows g, g + nblk, ...; output row
    nblk*r + g reads input rows nblk*r + (g + dy), so virtual tap (g + dy, dx) carries tap (dy, dx) in block g."""
    assert launch.si == 1 and launch.so == 1
    tapmap = dict(zip(launch.taps, launch.wtaps))
    dys = [dy for dy, _ in launch.taps]
    dxs = sorted({dx for _, dx in launch.taps})
    vt = [(j, dx) for j in range(min(dys), max(dys) + nblk) for dx in dxs]
    src = [[tapmap.get((j - g, dx)) for g in range(nblk)] for j, dx in vt]
    return Stacked(nblk, (launch.mi + nblk - 1) // nblk, launch.mj, nblk, nblk, 1,
                   [launch.oy0 + g for g in range(nblk)], [launch.ox0] * nblk, vt, src)
