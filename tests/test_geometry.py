"""CPU check of the tap tables (conv_geometry.py) against torch.nn.functional, through a torch emulation
of the gather primitive's definition in include/ast.h (no CUDA)."""
import pytest
import torch
import torch.nn.functional as F

from artist_style_transfer_b200 import conv_geometry as cg


def emulate_gather(x, wt, launches, out_hw, reflect=False):
    """x: [N,H,W,Ci]; wt: [T][Co][Ci]; returns [N,Ho,Wo,Co] following the formula in ast.h."""
    n, h, w, ci = x.shape
    co = wt.shape[1]
    out = torch.full((n, out_hw[0], out_hw[1], co), float("nan"), dtype=x.dtype)
    for l in launches:
        acc = torch.zeros((n, l.mi, l.mj, co), dtype=x.dtype)
        ii = torch.arange(l.mi)
        jj = torch.arange(l.mj)
        for t, (dy, dx) in enumerate(l.taps):
            y = l.si * ii + dy
            xx = l.si * jj + dx
            if reflect:
                y = y.abs(); y = torch.where(y >= h, 2 * (h - 1) - y, y)
                xx = xx.abs(); xx = torch.where(xx >= w, 2 * (w - 1) - xx, xx)
                oky = torch.ones_like(y, dtype=torch.bool); okx = torch.ones_like(xx, dtype=torch.bool)
            else:
                oky = (y >= 0) & (y < h); okx = (xx >= 0) & (xx < w)
            g = x[:, y.clamp(0, h - 1)][:, :, xx.clamp(0, w - 1)]
            g = g * (oky[:, None] & okx[None, :])[None, :, :, None]
            acc += g @ wt[l.woff + t].T
        oy = l.oy0 + l.so * ii
        ox = l.ox0 + l.so * jj
        ky, kx = oy < out_hw[0], ox < out_hw[1]
        out[:, oy[ky][:, None], ox[kx][None, :]] = acc[:, ky][:, :, kx]
    assert not torch.isnan(out).any(), "launches do not cover the output"
    return out


def pack(w, launches, a_dim, b_dim):
    """[T][a][b] from a 4-D weight whose dims (a_dim, b_dim, 2, 3) = (a, b, u, v)."""
    return torch.stack([w.select(3, v).select(2, u).permute(*((0, 1) if a_dim < b_dim else (1, 0)))
                        for u, v in cg.all_wtaps(launches)])


@pytest.mark.parametrize("k,s,h,w", [(9, 1, 12, 14), (3, 1, 8, 8), (3, 2, 10, 12), (1, 1, 5, 6), (3, 2, 9, 11)])
def test_conv_fwd_reflect_padded(k, s, h, w):
    torch.manual_seed(0)
    x = torch.randn(2, 5, h, w, dtype=torch.float64)
    wt = torch.randn(7, 5, k, k, dtype=torch.float64)
    xp = F.pad(x, (k // 2,) * 4, mode="reflect") if k > 1 else x
    ref = F.conv2d(xp, wt, stride=s)
    ls = cg.conv_fwd(k, s, 0, xp.shape[2], xp.shape[3])
    got = emulate_gather(xp.permute(0, 2, 3, 1), pack(wt, ls, 0, 1), ls, ref.shape[2:])
    torch.testing.assert_close(got.permute(0, 3, 1, 2), ref)
    # same thing from the unpadded image with the REFLECT flag
    ls2 = cg.conv_fwd(k, s, k // 2, h, w)
    got2 = emulate_gather(x.permute(0, 2, 3, 1), pack(wt, ls2, 0, 1), ls2, ref.shape[2:], reflect=True)
    torch.testing.assert_close(got2.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("k,s,pad,h,w", [(3, 1, 1, 8, 9), (3, 1, 0, 10, 10), (3, 2, 0, 10, 12), (9, 1, 0, 20, 20),
                                         (1, 1, 0, 4, 4), (3, 2, 0, 11, 9)])
def test_conv_dgrad(k, s, pad, h, w):
    torch.manual_seed(1)
    x = torch.randn(2, 4, h, w, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(6, 4, k, k, dtype=torch.float64)
    y = F.conv2d(x, wt, stride=s, padding=pad)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    ls = cg.conv_dgrad(k, s, pad, h, w)
    got = emulate_gather(gy.permute(0, 2, 3, 1), pack(wt, ls, 1, 0), ls, (h, w))   # [t][ci][co]
    torch.testing.assert_close(got.permute(0, 3, 1, 2), gx)


@pytest.mark.parametrize("k,s,op,h,w", [(3, 2, 1, 6, 7), (1, 1, 0, 5, 5)])
def test_convT_fwd_and_dgrad(k, s, op, h, w):
    torch.manual_seed(2)
    x = torch.randn(2, 4, h, w, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(4, 6, k, k, dtype=torch.float64)       # ConvTranspose2d layout (Ci,Co,k,k)
    y = F.conv_transpose2d(x, wt, stride=s, padding=k // 2, output_padding=op)
    ls = cg.convT_fwd(k, s, k // 2, op, h, w)
    got = emulate_gather(x.detach().permute(0, 2, 3, 1), pack(wt, ls, 1, 0), ls, y.shape[2:])   # [t][co][ci]
    torch.testing.assert_close(got.permute(0, 3, 1, 2), y.detach())
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    ld = cg.convT_dgrad(k, s, k // 2, y.shape[2], y.shape[3])
    gotd = emulate_gather(gy.permute(0, 2, 3, 1), pack(wt, ld, 0, 1), ld, (h, w))               # [t][ci][co]
    torch.testing.assert_close(gotd.permute(0, 3, 1, 2), gx)


def emulate_stacked(x, ws, stk, out, cb):
    """The formula of ast_conv_stacked (include/ast.h): x [N,H,W,Ci], ws [nvt][128][Ci]; fills the blocks' pixels of out."""
    n, h, w, ci = x.shape
    ii, jj = torch.arange(stk.mi), torch.arange(stk.mj)
    acc = torch.zeros((n, stk.mi, stk.mj, 128), dtype=x.dtype)
    for v, (dy, dx) in enumerate(stk.vt):
        y, xx = stk.sy * ii + dy, jj + dx
        ok = ((y >= 0) & (y < h))[:, None] & ((xx >= 0) & (xx < w))[None, :]
        g = x[:, y.clamp(0, h - 1)][:, :, xx.clamp(0, w - 1)] * ok[None, :, :, None]
        acc += g @ ws[v].T
    for b in range(stk.nblk):
        oy, ox = stk.oy[b] + stk.soy * ii, stk.ox[b] + stk.sox * jj
        ky, kx = oy < out.shape[1], ox < out.shape[2]
        out[:, oy[ky][:, None], ox[kx][None, :]] = acc[:, ky][:, :, kx][..., b * cb:(b + 1) * cb]


def _stack_filter(wt, launches, stk, cb):
    tidx = {}
    for l in launches:
        for t, pos in enumerate(l.wtaps):
            tidx[pos] = l.woff + t
    ws = torch.zeros(len(stk.vt), 128, wt.shape[2], dtype=wt.dtype)
    for v, row in enumerate(stk.src):
        for g, pos in enumerate(row):
            if pos is not None:
                ws[v, g * cb:(g + 1) * cb] = wt[tidx[pos]]
    return ws


@pytest.mark.parametrize("case", ["convT32", "dgrad32", "convT64", "rows9", "rows3_64", "rows9_neg", "rows3x3"])
def test_stacked_launches_equal_the_plain_gather(case):
    """stack_phases / stack_rows (the block-stacked tensor-core kernel's geometry) reproduce the plain launches."""
    torch.manual_seed(1)
    cb = 64 if "64" in case else 32
    if case in ("convT32", "convT64"):
        hin, win, ls = 7, 9, cg.convT_fwd(3, 2, 1, 1, 7, 9)
        ho, wo = 14, 18
    elif case == "dgrad32":
        hin, win, ho, wo = 6, 7, 14, 16          # data gradient of a stride-2 3x3 conv on a padded 14 x 16 input
        ls = cg.conv_dgrad(3, 2, 0, ho, wo)
    elif case == "rows9":
        hin, win, ho, wo = 19, 10, 11, 10
        ls = [cg.Launch(ho, wo, 1, 1, 0, 0, [(d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
    elif case == "rows9_neg":
        hin, win, ho, wo = 10, 9, 18, 9           # thin-output data gradient: rows y - d, zero outside
        ls = [cg.Launch(ho, wo, 1, 1, 0, 0, [(-d, 0) for d in range(9)], [(d, 0) for d in range(9)], 0)]
    elif case == "rows3_64":
        hin, win, ho, wo = 9, 8, 9, 8
        ls = [cg.Launch(ho, wo, 1, 1, 0, 0, [(-1, 0), (0, 0), (1, 0)], [(0, 0), (1, 0), (2, 0)], 0)]
    else:
        hin, win, ho, wo = 10, 11, 10, 11
        ls = cg.conv_fwd(3, 1, 1, hin, win)
    nt = sum(len(l.taps) for l in ls)
    x = torch.randn(2, hin, win, 6, dtype=torch.float64)
    wt = torch.randn(nt, cb, 6, dtype=torch.float64)
    ref = emulate_gather(x, wt, ls, (ho, wo))
    nb = 128 // cb
    got = torch.full_like(ref, float("nan"))
    if len(ls) > 1:           # sub-pixel phases: one launch of 4 blocks (32 channels) or two launches of 2 (64 channels)
        groups = [cg.stack_phases(ls[i:i + nb]) for i in range(0, len(ls), nb)]
        assert [len(s.vt) for s in groups] == ([4] if nb == 4 else [2, 4])
        assert sum(s.ntaps for s in groups) == nt
    else:                     # interleaved rows: k + nblk - 1 virtual rows, every tap once per block
        groups = [cg.stack_rows(ls[0], nb)]
        assert len(groups[0].vt) == (len({dy for dy, _ in ls[0].taps}) + nb - 1) * len({dx for _, dx in ls[0].taps})
        assert groups[0].ntaps == nt * nb
    for stk in groups:
        emulate_stacked(x, _stack_filter(wt, ls, stk, cb), stk, got, cb)
    assert not torch.isnan(got).any(), "blocks do not cover the output"
    torch.testing.assert_close(got, ref)


def test_sizes():
    assert cg.conv_out_size(258, 3, 2, 0) == 128
    assert cg.convT_out_size(64, 3, 2, 1, 1) == 128
    assert sum(len(l.taps) for l in cg.convT_fwd(3, 2, 1, 1, 8, 8)) == 9
    assert [len(l.taps) for l in cg.convT_fwd(3, 2, 1, 1, 8, 8)] == [1, 2, 2, 4]      # SURVEY 8c sub-pixel phases


def _emulate_pack(w, pack, dims):
    """What ast_adam_step writes: element (a, b, u, v) -> off + a*stride[0] + b*stride[1] + taps[u*V + v] (+ r*rep_stride)."""
    A, B, U, V = dims
    n = 1
    for d in pack.shape:
        n *= d
    ia = torch.arange(A).view(A, 1, 1) * pack.stride[0]
    ib = torch.arange(B).view(1, B, 1) * pack.stride[1]
    it = torch.tensor(pack.taps, dtype=torch.int64).view(1, 1, U * V)
    offs = (ia + ib + it).reshape(-1)
    flat = torch.zeros(n, dtype=torch.float64)
    count = torch.zeros(n, dtype=torch.int64)
    for r in range(pack.rep):
        count += torch.bincount(offs + r * pack.rep_stride, minlength=n)
        flat[offs + r * pack.rep_stride] = w.reshape(-1)
    assert int(count.max()) == 1, "two master elements map to the same packed slot"
    return flat.view(pack.shape)


def _check_stacked(got, pack, cb, tile_of):
    """A stacked filter holds tile_of(kernel position) in rows [b*cb, (b+1)*cb) of virtual tap v wherever the canonical
    conv_geometry.Stacked launch lists that position, and zeros everywhere else."""
    seen = 0
    for grp, vb in zip(pack.stacked, pack.vbase):
        for v, row in enumerate(grp.src):
            for b, pos in enumerate(row):
                blk = got[vb + v, b * cb:(b + 1) * cb]
                if pos is None:
                    assert float(blk.abs().sum()) == 0.0
                else:
                    assert torch.equal(blk, tile_of(pos)), (v, b, pos)
                    seen += 1
    assert seen == sum(g.ntaps for g in pack.stacked)


@pytest.mark.parametrize("mode", ["fp32", "fast"])
def test_arena_layouts_match_the_gather_kernels_operand_layouts(mode):
    """arena.TransferArena (host logic, CPU): the table-driven pack maps reproduce the [tap][cout][cin] operand layouts
    that conv_geometry's launches index, for conv / stride-2 conv / ConvTranspose phases / the 3-channel ends; gradient
    segments are disjoint and 16-byte aligned and the strided p.grad views address the same elements as the maps."""
    import torch
    from artist_style_transfer_b200 import arena as arena_mod, cnn
    net = cnn.StyleTransfer(device="cpu", precision=mode)
    stages = net._stages()
    ar = arena_mod.TransferArena(stages, mode, torch.device("cpu"))
    spans = []
    for st, pl in zip(stages, ar.plans):
        w = st.conv.weight.detach().double()
        k, co, ci = st.k, st.cout, st.cin
        dims = tuple(w.shape)
        conv = st.kind == "conv"
        # ---- forward pack
        got = _emulate_pack(w, pl.fwd, dims)
        if pl.fwd.stacked is not None:
            assert mode == "fast"
            if pl.thin_in:
                def tile(pos):
                    t = torch.zeros(co, 32, dtype=torch.float64)
                    for dx in range(k):
                        t[:, dx * ci:(dx + 1) * ci] = w[:, :, pos[0], dx]
                    return t
                _check_stacked(got, pl.fwd, co, tile)
            elif pl.thin_out:
                def tile(pos):
                    t = torch.zeros(32, ci, dtype=torch.float64)
                    for dx in range(k):
                        t[dx * co:(dx + 1) * co] = w[:, :, pos[0], dx]
                    return t
                _check_stacked(got, pl.fwd, 32, tile)
            else:
                assert not conv and st.stride == 2
                _check_stacked(got, pl.fwd, co, lambda pos: w[:, :, pos[0], pos[1]].t())
        elif pl.thin_in:
            for dy in range(k):
                for dx in range(k):
                    assert torch.equal(got[dy, :, dx * ci:(dx + 1) * ci], w[:, :, dy, dx])
            assert float(got[:, :, k * ci:].abs().sum()) == 0.0
        elif pl.thin_out:
            for dy in range(k):
                for dx in range(k):
                    assert torch.equal(got[dy, dx * co:(dx + 1) * co, :], w[:, :, dy, dx])
        else:
            launches = cnn._fwd_geometry(st, 16 + k, 16 + k)[0]
            ar.woff(pl.fwd, launches)
            for l in launches:
                for t, (u, v) in enumerate(l.wtaps):
                    ref = w[:, :, u, v] if conv else w[:, :, u, v].t()           # [co][ci]
                    assert torch.equal(got[l.woff + t], ref), (st.kind, k, st.stride, u, v)
        # ---- data-gradient pack
        if pl.dgrad is not None:
            got = _emulate_pack(w, pl.dgrad, dims)
            if pl.dgrad.stacked is not None:
                if pl.thin_out:
                    def tile(pos):
                        t = torch.zeros(ci, 32, dtype=torch.float64)
                        for dx in range(k):
                            t[:, dx * co:(dx + 1) * co] = w[:, :, pos[0], dx].t()
                        return t
                    _check_stacked(got, pl.dgrad, ci, tile)
                else:
                    assert conv and st.stride == 2
                    _check_stacked(got, pl.dgrad, ci, lambda pos: w[:, :, pos[0], pos[1]].t())
            elif pl.thin_out:
                for dy in range(k):
                    for dx in range(k):
                        assert torch.equal(got[dy, :, dx * co:(dx + 1) * co], w[:, :, dy, dx].t())
            else:
                dl = (cg.conv_dgrad(k, st.stride, 0, 16 + k, 16 + k) if conv else cg.convT_dgrad(k, st.stride, k // 2, 16, 16))
                ar.woff(pl.dgrad, dl)
                for l in dl:
                    for t, (u, v) in enumerate(l.wtaps):
                        ref = w[:, :, u, v].t() if conv else w[:, :, u, v]       # [ci][co]
                        assert torch.equal(got[l.woff + t], ref)
        n = 1
        for d in pl.g_shape:
            n *= d
        spans += [(pl.g_off, pl.g_off + n), (pl.g_cb, pl.g_cb + co)]
        if st.norm:
            spans += [(pl.g_gam, pl.g_gam + co), (pl.g_bet, pl.g_bet + co)]
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0 and b0 % 4 == 0
    assert spans[-1][1] <= ar.g_numel
    # strided gradient views == the (a, b, tap) maps of the descriptors
    gbuf = torch.arange(ar.g_numel, dtype=torch.float32)
    views = ar.grad_views(gbuf)
    it = iter(views)
    for st, pl in zip(stages, ar.plans):
        gw = next(it)
        assert tuple(gw.shape) == tuple(st.conv.weight.shape)
        k = st.k
        for (a, b, u, v) in [(0, 0, 0, 0), (1, 2, k - 1, 0), (2, 1, 0, k - 1), (gw.shape[0] - 1, gw.shape[1] - 1, k - 1, k - 1)]:
            want = pl.g_off + a * pl.g_stride[0] + b * pl.g_stride[1] + pl.g_taps[u * k + v]
            assert int(gw[a, b, u, v]) == want
        assert int(next(it)[0]) == pl.g_cb
        if st.norm:
            assert int(next(it)[0]) == pl.g_gam and int(next(it)[0]) == pl.g_bet


@pytest.mark.parametrize("ntaps,h", [(9, 11), (9, 4), (5, 7), (12, 6)])
def test_merged_thin_filter_gradient_blocks(ntaps, h):
    """Algebra behind contract_thin's merged mode (contract_tc.cu): for a vertical tap stack
        dW[t][cc][rc] = sum_r X[r + t][cc] * G[r][rc]
    ONE M128 x N(32*ngrp) product per row chunk i0, A = X rows i0 .. i0+3 (block t1), B = G rows i0 - 4*(ngrp-1-nb)
    (block nb), accumulated over i0 = 0 .. h + 4*(ngrp-1) - 1 with rows outside the images reading 0, leaves tap
    t = t1 + 4*(ngrp-1-nb) in block (t1, nb)."""
    torch.manual_seed(ntaps * h)
    w, cc, rc = 5, 3, 4
    ngrp = (ntaps + 3) // 4
    X = torch.randn(h + ntaps - 1, w, cc, dtype=torch.float64)     # the shifted operand (e.g. the padded layer input)
    G = torch.randn(h, w, rc, dtype=torch.float64)                 # the fixed operand (the output gradient)
    ref = torch.stack([torch.einsum("rwc,rwd->cd", X[t:t + h], G) for t in range(ntaps)])

    def row(img, r):
        return img[r] if 0 <= r < img.shape[0] else torch.zeros_like(img[0])

    acc = torch.zeros(4, cc, ngrp, rc, dtype=torch.float64)
    for i0 in range(h + 4 * (ngrp - 1)):
        for t1 in range(4):
            for nb in range(ngrp):
                acc[t1, :, nb] += torch.einsum("wc,wd->cd", row(X, i0 + t1), row(G, i0 - 4 * (ngrp - 1 - nb)))
    for t in range(ntaps):
        t1, t2 = t % 4, t // 4
        torch.testing.assert_close(acc[t1, :, ngrp - 1 - t2], ref[t])
