"""Module- and step-level parity of the CUDA path against the CPU oracle and the committed golden vectors
(generated from the reference classes).  north_star tolerances: strict fp32 1e-5 on Grams and losses;
bf16/TF32 mode 2e-3 on Grams, 1e-2 on losses."""
import os

import numpy as np
import pytest
import torch

from oracle import port, weights

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ast():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import artist_style_transfer_b200 as m
    return m


def build(ast, precision, seed=2):
    net = ast.StyleTransfer(device=torch.device("cuda"), precision=precision)
    net.load_state_dict(weights.transfer_state_dict(seed), strict=True)
    vgg = ast.VGG16(vgg_path=None, precision=precision).cuda()
    missing, unexpected = vgg.load_state_dict(weights.vgg_state_dict(seed), strict=False)
    assert not missing and not unexpected
    return net, vgg


def test_transfer_forward_fp32(ast):
    net, _ = build(ast, "fp32")
    x = weights.content_batch(2, 64, 2)
    y = net(x.cuda())
    yr = port.transfer_forward(x.double(), {k: v.double() for k, v in weights.transfer_state_dict(2).items()})
    assert y.shape == yr.shape
    assert rel(y, yr) < 1e-5


def test_transfer_forward_nonsquare_and_batch1(ast):
    net, _ = build(ast, "fp32")
    x = weights.content_batch(1, 40, 2, width=56)
    y = net(x.cuda())
    yr = port.transfer_forward(x.double(), {k: v.double() for k, v in weights.transfer_state_dict(2).items()})
    assert rel(y, yr) < 1e-5


def test_vgg_forward_fp32(ast):
    _, vgg = build(ast, "fp32")
    x = weights.content_batch(2, 64, 2)
    nm = port.neg_mean()
    feats = vgg((x + nm).cuda())
    fr = port.vgg_features((x + nm).double(), {k: v.double() for k, v in weights.vgg_state_dict(2).items()})
    assert list(feats.keys()) == ["relu1_2", "relu2_2", "relu3_3", "relu4_3"]
    for k in fr:
        assert feats[k].shape == fr[k].shape
        assert rel(feats[k], fr[k]) < 1e-5, k
    fused = vgg(x.cuda(), shift=ast.neg_mean(torch.device("cuda")))          # mean shift fused in the conv1_1 loader
    for k in fr:
        assert rel(fused[k], fr[k]) < 1e-5, k
    just = ast.VGG16(just_content=True, vgg_path=None, precision="fp32").cuda()
    just.load_state_dict(weights.vgg_state_dict(2), strict=False)
    assert rel(just((x + nm).cuda()), fr["relu2_2"]) < 1e-5


def _step(ast, precision, batch, size):
    net, vgg = build(ast, precision)
    content = weights.content_batch(batch, size, 2).cuda()
    style = ast.style_grams_single(vgg, weights.style_image(size, 2).cuda(), batch)
    net.zero_grad()
    c, s, t = ast.perceptual_step(net, vgg, content, style)
    with torch.no_grad():
        gen = net(content)
        grams = {k: ast.gram(v) for k, v in vgg(gen, shift=ast.neg_mean(content.device)).items()}
    return net, (float(c), float(s), float(t)), grams, style


def _gram_checks(grams, gold, tol, prefix="gram"):
    for k, g in grams.items():
        g = g.double().cpu()
        ref = gold[f"{prefix}/{k}/block"]
        err = np.linalg.norm(g[:, :32, :32].numpy() - ref) / np.linalg.norm(ref)
        assert err < tol, (k, err)
        np.testing.assert_allclose(g.flatten(1).norm(dim=1).numpy(), gold[f"{prefix}/{k}/fro"], rtol=tol)
        np.testing.assert_allclose(torch.diagonal(g, dim1=1, dim2=2).numpy(), gold[f"{prefix}/{k}/diag"],
                                   rtol=10 * tol, atol=tol * float(np.abs(gold[f"{prefix}/{k}/diag"]).max()))


def test_step_strict_vs_golden_small(ast, golden_dir):
    """B=2, 64^2, strict mode vs the reference's own fp64 outputs, incl. parameter gradients."""
    gold = np.load(os.path.join(golden_dir, "step_b2_s64_f64.npz"))
    net, losses, grams, style = _step(ast, "fp32", 2, 64)
    np.testing.assert_allclose(np.array(losses), gold["losses"], rtol=1e-5)
    _gram_checks(grams, gold, 1e-5)
    _gram_checks(style, gold, 1e-5, prefix="style_gram")
    # gradients: fp32-vs-fp64 noise floor through 17 InstanceNorms is ~1e-2 at 256^2; at 64^2 we
    # measure norms to 2e-3 here and pin exact gradients per layer in test_layers_gpu.py
    for name, p in net.named_parameters():
        ref = float(gold["grad_norm/" + name])
        if ref < 1e-9:
            assert float(p.grad.norm()) < 1e-6
            continue
        assert abs(float(p.grad.norm()) - ref) < 5e-3 * ref, (name, float(p.grad.norm()), ref)


def test_step_strict_config1(ast, golden_dir):
    """BASELINE config 1: B=4, 256^2, strict fp32 within 1e-5 of the reference (fp64 and fp32 runs)."""
    g64 = np.load(os.path.join(golden_dir, "step_b4_s256_f64.npz"))
    _, losses, grams, _ = _step(ast, "fp32", 4, 256)
    np.testing.assert_allclose(np.array(losses), g64["losses"], rtol=1e-5)
    _gram_checks(grams, g64, 1e-5)


def test_step_fast_config1(ast, golden_dir):
    """bf16/TF32-input mode at config 1: 2e-3 on Grams, 1e-2 on losses."""
    g64 = np.load(os.path.join(golden_dir, "step_b4_s256_f64.npz"))
    _, losses, grams, _ = _step(ast, "fast", 4, 256)
    np.testing.assert_allclose(np.array(losses), g64["losses"], rtol=1e-2)
    _gram_checks(grams, g64, 2e-3)


def test_step_grads_vs_oracle_fp32(ast):
    """Same inputs through oracle (fp64) and CUDA strict path at 32^2: every parameter gradient."""
    net, vgg = build(ast, "fp32")
    content = weights.content_batch(2, 32, 2)
    style_img = weights.style_image(32, 2)
    tsd = port.make_leaf(weights.transfer_state_dict(2), torch.float64)
    vsd = {k: v.double() for k, v in weights.vgg_state_dict(2).items()}
    sg = port.style_grams_single(style_img.double(), vsd, 2)
    ref = port.training_step(tsd, vsd, content.double(), sg)
    style = ast.style_grams_single(vgg, style_img.cuda(), 2)
    net.zero_grad()
    c, s, t = ast.perceptual_step(net, vgg, content.cuda(), style)
    assert abs(float(t) - float(ref["total"])) < 1e-5 * float(ref["total"])
    worst = 0.0
    for name, p in net.named_parameters():
        gr = ref["grads"][name]
        if float(gr.norm()) < 1e-9:
            continue
        worst = max(worst, rel(p.grad, gr))
    assert worst < 2e-3, worst


def test_smartaverage_vs_golden(ast, golden_dir):
    gold = np.load(os.path.join(golden_dir, "smartavg_b2_s64_n5_f64.npz"))
    _, vgg = build(ast, "fp32")
    paintings = [weights.style_image(64, 2, i).cuda() for i in range(5)]
    grams = ast.style_grams_smartaverage(vgg, paintings, 2, mode="reference")
    _gram_checks(grams, gold, 1e-5)
    assert grams["relu1_2"].shape == (2, 64, 64)
    alt = ast.style_grams_smartaverage(vgg, paintings, 2, mode="mean_gram")
    vsd = {k: v.double() for k, v in weights.vgg_state_dict(2).items()}
    ref_alt = port.style_grams_smartaverage([p.double().cpu() for p in paintings], vsd, 2, mode="mean_gram")
    for k in alt:
        assert rel(alt[k], ref_alt[k]) < 1e-5, k


def test_trainer_step_matches_oracle_adam(ast):
    """Two optimizer steps (Adam + L2, train_cnn.py:247,334) track the oracle's parameters."""
    net, vgg = build(ast, "fp32")
    style_img = weights.style_image(32, 2)
    style = ast.style_grams_single(vgg, style_img.cuda(), 2)
    trainer = ast.PerceptualTrainer(net, vgg, style, lr=1e-3)
    vsd = {k: v.double() for k, v in weights.vgg_state_dict(2).items()}
    sg = port.style_grams_single(style_img.double(), vsd, 2)
    params = {k: v.double() for k, v in weights.transfer_state_dict(2).items()}
    state = {}
    for step in (1, 2):
        content = weights.content_batch(2, 32, 2, step=step)
        trainer.step(content.cuda())
        leaf = port.make_leaf(params, torch.float64)
        out = port.training_step(leaf, vsd, content.double(), sg)
        params = port.adam_l2_step({k: v.detach() for k, v in leaf.items()}, out["grads"], state, 1e-3, step)
    for name, p in net.named_parameters():
        if name.endswith("conv_layer.bias") or name.endswith("conv_transpose.bias"):
            continue
        # Adam's first steps move every weight by ~lr*sign(g): fp32-vs-fp64 noise on near-zero gradients flips a few
        # signs, so compare the mean displacement against lr rather than element-wise
        diff = (p.detach().double().cpu() - params[name]).abs()
        assert float(diff.mean()) < 0.05 * 1e-3, (name, float(diff.mean()))
        assert float(diff.max()) < 2.5 * 2 * 1e-3, (name, float(diff.max()))


def test_launch_counter(ast):
    from artist_style_transfer_b200 import _lib
    before = _lib.launch_count()
    ast.gram(torch.randn(1, 64, 8, 8, device="cuda"))
    assert _lib.launch_count() > before


def test_fast_mode_runs_on_tensor_kernels_only(ast):
    """In fast mode every conv / wgrad / Gram launch of a step is a tcgen05 kernel (no SIMT conv left)."""
    from artist_style_transfer_b200 import ops
    net, vgg = build(ast, "fast")
    content = weights.content_batch(2, 64, 2).cuda()
    style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
    ast.perceptual_step(net, vgg, content, style)
    net.zero_grad()
    ops.profile_begin()
    ast.perceptual_step(net, vgg, content, style)
    fam = ops.profile_end()
    assert "conv_gather_tc" in fam and "wgrad_tc" in fam
    assert "conv_gather_simt" not in fam and "wgrad_simt" not in fam, fam


def test_fast_vs_strict_gradients(ast):
    """bf16/TF32 tensor-core step against the strict fp32 step on identical inputs: losses 1e-2, gradient direction."""
    content = weights.content_batch(2, 64, 2).cuda()
    out = {}
    for mode in ("fp32", "fast"):
        net, vgg = build(ast, mode)
        style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
        net.zero_grad()
        c, s, t = ast.perceptual_step(net, vgg, content, style)
        out[mode] = (float(t), {n: p.grad.clone() for n, p in net.named_parameters()})
    assert abs(out["fast"][0] - out["fp32"][0]) < 1e-2 * abs(out["fp32"][0])
    for name, g in out["fp32"][1].items():
        if float(g.norm()) < 1e-9 or name.endswith("conv2.norm_layer.bias"):   # loss-invariant parameters: |g| ~ 0
            continue
        gf = out["fast"][1][name]
        cos = float((g * gf).sum() / (g.norm() * gf.norm() + 1e-30))
        # bf16 activations perturb the forward point; the deviation accumulates towards the input side (measured
        # 0.945..1.0, and 0.97 between two bf16 runs that differ only in summation order): see DESIGN.md
        assert cos > 0.9, (name, cos)
        assert abs(float(gf.norm()) / float(g.norm()) - 1) < 0.15, name


def test_cuda_graph_step_matches_eager(ast):
    """Capturing the whole step in a CUDA graph must not change the numbers (same kernels, same order)."""
    content = [weights.content_batch(2, 64, 2, step=i).cuda() for i in range(6)]
    res = {}
    for graph in (False, True):
        net, vgg = build(ast, "fast")
        style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
        tr = ast.PerceptualTrainer(net, vgg, style, lr=1e-3, cuda_graph=graph)
        losses = [tuple(float(v) for v in tr.step(c)) for c in content]
        res[graph] = (losses, [p.detach().clone() for p in net.parameters()])
    for a, b in zip(res[False][0], res[True][0]):
        np.testing.assert_allclose(np.array(a), np.array(b), rtol=2e-3)     # atomics order differs run to run
    assert res[True][0][3] != res[True][0][4]                               # replays really consume new inputs
    # Parameters: Adam moves every weight by ~lr*sign(g) per step and bf16 run-to-run noise flips near-zero gradients,
    # so two runs of the SAME code drift apart by a fraction of the 6*lr a weight can travel; only require that the graph
    # run stays within that envelope and is finite (the per-step losses above are the tight check).
    for pa, pb in zip(res[False][1], res[True][1]):
        assert torch.isfinite(pb).all()
        assert float((pa - pb).abs().max()) <= 2 * 6 * 1e-3 + 1e-6


@pytest.mark.parametrize("graph", [False, True])
def test_prefetch_pipeline_matches_direct_steps(ast, graph):
    """PerceptualTrainer.prefetch(): pinned host batches copied on a side stream one step ahead give the same losses as
    handing device tensors to step() (the overlap must not reorder or drop a batch)."""
    host = [weights.content_batch(2, 64, 2, step=i).pin_memory() for i in range(6)]
    res = {}
    for mode in ("direct", "prefetch"):
        net, vgg = build(ast, "fast")
        style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
        tr = ast.PerceptualTrainer(net, vgg, style, lr=1e-3, cuda_graph=graph)
        losses = []
        if mode == "direct":
            for h in host:
                losses.append(tuple(float(v) for v in tr.step(h.cuda())))
        else:
            nxt = tr.prefetch(host[0])
            for i in range(len(host)):
                cur = nxt
                if i + 1 < len(host):
                    nxt = tr.prefetch(host[i + 1])
                losses.append(tuple(float(v) for v in tr.step(cur)))
        res[mode] = losses
    for a, b in zip(res["direct"], res["prefetch"]):
        np.testing.assert_allclose(np.array(a), np.array(b), rtol=2e-3)
    assert res["prefetch"][3] != res["prefetch"][4]
