"""Fused flat-arena Adam(L2) + StepLR + weight re-pack (train_cnn.py:247-248,334,375) against the oracle's
`adam_l2_step`, torch.optim.Adam and the per-layer pack kernels."""
import numpy as np
import pytest
import torch

from oracle import port, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ast():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import artist_style_transfer_b200 as m
    return m


def _net(ast, precision="fast"):
    net = ast.StyleTransfer(device=torch.device("cuda"), precision=precision)
    net.load_state_dict(weights.transfer_state_dict(2), strict=True)
    return net


@pytest.mark.parametrize("precision", ["fast", "fp32"])
def test_fused_adam_matches_oracle_and_repacks(ast, precision):
    """Three updates with known gradients written through the strided p.grad views: parameters within 1e-6 of
    port.adam_l2_step (fp64) per step, packed operand copies == a fresh per-layer pack of the updated parameters."""
    from artist_style_transfer_b200 import cnn, ops
    net = _net(ast, precision)
    arena = net._arena_for(torch.device("cuda", torch.cuda.current_device()))
    arena.enable_optimizer(lr=2e-3, weight_decay=1e-4)
    gbuf = arena.new_grad_buffer()
    views = arena.grad_views(gbuf)
    params = arena.params()
    names = {id(p): n for n, p in net.named_parameters()}
    ref = {names[id(p)]: p.detach().double().cpu() for p in params}
    state = {}
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in (1, 2, 3):
        gbuf.zero_()
        grads = {}
        for p, v in zip(params, views):
            r = torch.randn(p.shape, device="cuda", generator=g) * 0.1
            v.copy_(r)                                     # strided write into the tap-major arena
            grads[names[id(p)]] = r.double().cpu()
        if step == 3:
            arena.set_lr(1e-3)                             # StepLR halving: read from device memory by the kernel
        arena.adam_step(gbuf)
        ref = port.adam_l2_step(ref, grads, state, 2e-3 if step < 3 else 1e-3, step)
        for p in params:
            want = ref[names[id(p)]]
            err = float((p.detach().double().cpu() - want).abs().max())
            assert err < 1e-6 * max(1.0, float(want.abs().max())), (names[id(p)], step, err)
    # packs written by the Adam kernel == per-layer pack kernels on the updated parameters
    adt = torch.bfloat16 if precision == "fast" else torch.float32
    for st, pl in zip(arena.stages, arena.plans):
        if pl.thin_in or pl.thin_out:
            continue
        launches = cnn._fwd_geometry(st, 16 + st.k, 16 + st.k)[0]
        k2 = st.k * st.k
        w = st.conv.weight.detach()
        if st.kind == "conv":
            want = ops.pack_weights(w, launches, st.cout, st.cin, st.cin * k2, k2, st.k, 1, adt)
        else:
            want = ops.pack_weights(w, launches, st.cout, st.cin, k2, st.cout * k2, st.k, 1, adt)
        if pl.fwd.stacked is not None:      # block-stacked filter (ConvTranspose phases): the same tiles, re-arranged
            tidx = {wt: l.woff + t for l in launches for t, wt in enumerate(l.wtaps)}
            for grp, vb in zip(pl.fwd.stacked, pl.fwd.vbase):
                ws = ops.stack_filter(lambda pos: want[tidx[pos]], grp, st.cout, st.cin, adt, want.device)
                assert torch.equal(pl.fwd.tensor[vb:vb + len(grp.vt)], ws), (st.kind, st.k, st.stride)
            continue
        assert torch.equal(pl.fwd.tensor, want), (st.kind, st.k, st.stride)


def test_fused_vs_torch_adam_one_step(ast):
    """Same gradients -> torch.optim.Adam(weight_decay) and the fused kernel move the parameters identically (1e-6)."""
    neta, netb = _net(ast), _net(ast)
    arena = neta._arena_for(torch.device("cuda", torch.cuda.current_device()))
    arena.enable_optimizer(lr=2.4e-3, weight_decay=1e-4)
    gbuf = arena.new_grad_buffer()
    opt = torch.optim.Adam(netb.parameters(), lr=2.4e-3, weight_decay=1e-4)
    pb = dict(netb.named_parameters())
    names = {id(p): n for n, p in neta.named_parameters()}
    g = torch.Generator(device="cuda").manual_seed(7)
    for _ in range(2):
        gbuf.zero_()
        for p, v in zip(arena.params(), arena.grad_views(gbuf)):
            r = torch.randn(p.shape, device="cuda", generator=g)
            v.copy_(r)
            pb[names[id(p)]].grad = r.clone()
        arena.adam_step(gbuf)
        opt.step()
    for p in arena.params():
        q = pb[names[id(p)]]
        assert float((p - q).abs().max()) < 1e-6, names[id(p)]


@pytest.mark.parametrize("optimizer", ["fused", "torch"])
def test_steplr_acts_under_cuda_graph(ast, optimizer):
    """with cuda_graph=True the lr must not be frozen into the captured graph.  step_size = 1 epoch:
    after end_epoch() the next replayed step moves the parameters by ~half as much, like the eager trainer."""
    content = [weights.content_batch(2, 64, 2, step=i).cuda() for i in range(8)]
    moves = {}
    for graph in (False, True):
        net = _net(ast)
        vgg = ast.VGG16(vgg_path=None, precision="fast").cuda()
        vgg.load_state_dict(weights.vgg_state_dict(2), strict=False)
        style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
        tr = ast.PerceptualTrainer(net, vgg, style, lr=1e-3, num_epochs=2, num_steps=2, cuda_graph=graph, optimizer=optimizer)
        w = net.ResidualBlock[2].conv1.conv_layer.weight
        for c in content[:5]:                 # 3 eager warm-up steps + capture + 1 replay when graph=True
            tr.step(c)
        torch.cuda.synchronize()
        before = w.detach().clone()
        tr.step(content[5])
        torch.cuda.synchronize()
        d_full = float((w.detach() - before).abs().mean())
        tr.end_epoch()                        # lr 1e-3 -> 5e-4
        assert abs(tr.lr - 5e-4) < 1e-9
        before = w.detach().clone()
        tr.step(content[6])
        torch.cuda.synchronize()
        d_half = float((w.detach() - before).abs().mean())
        moves[graph] = (d_full, d_half)
        tr.close()
    for graph, (d_full, d_half) in moves.items():
        # Adam's per-step displacement is ~lr: halving lr halves it (to within the step-to-step variation of m/sqrt(v))
        assert 0.35 < d_half / d_full < 0.65, (optimizer, graph, d_full, d_half)
    np.testing.assert_allclose(moves[True][1], moves[False][1], rtol=0.2)
