"""Parity at the BASELINE configs beyond config 1 (VERDICT r01 "untested configs"), against golden vectors generated
from the unmodified reference classes (oracle/make_golden.py):
  configs[1]  B=32 at 256^2, fast mode - the exact bench workload
  configs[2]  'smartaverage' over 512^2 paintings, fast mode, both semantics
  configs[3]  StyleTransfer.forward at 1080 x 1920 (non-power-of-two tiles, 270 x 480 bottleneck), both modes
  configs[4]  1024^2 step (relu1_2 Gram with K = 1,048,576 pixels), both modes
north_star tolerances: strict fp32 1e-5 on Grams / losses; bf16/TF32 mode 2e-3 on Grams, 1e-2 on losses."""
import os

import numpy as np
import pytest
import torch

from oracle import port, weights

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
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
    vgg.load_state_dict(weights.vgg_state_dict(seed), strict=False)
    return net, vgg


def gram_checks(grams, gold, tol, prefix="gram"):
    for k, g in grams.items():
        g = g.double().cpu()
        ref = gold[f"{prefix}/{k}/block"]
        err = np.linalg.norm(g[:, :32, :32].numpy() - ref) / np.linalg.norm(ref)
        assert err < tol, (k, err)
        np.testing.assert_allclose(g.flatten(1).norm(dim=1).numpy(), gold[f"{prefix}/{k}/fro"], rtol=tol)
        np.testing.assert_allclose(g.flatten(1).sum(dim=1).numpy(), gold[f"{prefix}/{k}/sum"], rtol=tol)


def losses_and_grams(ast, precision, batch, size):
    """Forward part of train_cnn.py:299-329 through the product path: losses from the fused loss node, Grams from it too."""
    net, vgg = build(ast, precision)
    content = weights.content_batch(batch, size, 2).cuda()
    style = ast.style_grams_single(vgg, weights.style_image(size, 2).cuda(), batch)
    shift = ast.neg_mean(content.device)
    with torch.no_grad():
        gen = net(content)
        cf = vgg(content, shift=shift, upto="relu2_2", only_last=True)["relu2_2"]
        gf = vgg(gen, shift=shift)
        c, s, grams = ast.perceptual_losses(gf, cf, style, fast=precision == "fast")
    return (float(c), float(s), float(c) + float(s)), grams, style, gen


def test_config2_bench_workload_b32_fast(ast, golden_dir):
    """B=32 at 256^2 in fast mode (what bench.py times) vs the reference's fp32 run of the same 32 images."""
    gold = np.load(os.path.join(golden_dir, "step_b32_s256_f32.npz"))
    losses, grams, style, _ = losses_and_grams(ast, "fast", 32, 256)
    np.testing.assert_allclose(np.array(losses), gold["losses"], rtol=1e-2)
    gram_checks(grams, gold, 2e-3)
    gram_checks({k: v.contiguous() for k, v in style.items()}, gold, 2e-3, prefix="style_gram")
    # and through the full training step (CUDA path incl. backward), same losses
    net, vgg = build(ast, "fast")
    c, s, t = ast.perceptual_step(net, vgg, weights.content_batch(32, 256, 2).cuda(), style)
    np.testing.assert_allclose(np.array([float(c), float(s), float(t)]), gold["losses"], rtol=1e-2)
    assert all(torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("precision,ltol,gtol", [("fp32", 2e-5, 2e-5), ("fast", 1e-2, 2e-3)])
def test_config5_1024_step(ast, golden_dir, precision, ltol, gtol):
    """1024^2, B=1: relu1_2 Gram reduces over 1,048,576 pixels (split-K with 16384 chunks), relu4_3 is 512 x 128^2.
    The golden is the reference in fp32 (its own noise floor vs fp64 is ~2e-7 on Grams), hence 2e-5 in strict mode."""
    gold = np.load(os.path.join(golden_dir, "step_b1_s1024_f32.npz"))
    losses, grams, style, _ = losses_and_grams(ast, precision, 1, 1024)
    np.testing.assert_allclose(np.array(losses), gold["losses"], rtol=ltol)
    gram_checks(grams, gold, gtol)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("fast", 5e-2)])
def test_config4_forward_1080p(ast, golden_dir, precision, tol):
    """StyleTransfer.forward at 1080 x 1920 (cnn.py:45-49, inference.py:115) vs the reference in fp64.  Strict: 1e-5.
    Fast mode runs bf16 activations through 17 InstanceNorms: the generated IMAGE is only accurate to ~2e-2 relative
    (SURVEY 8c'; north_star states its fast tolerances on Grams and losses, not on pixels) - 5e-2 asserted here."""
    gold = np.load(os.path.join(golden_dir, "fwd_b1_1080x1920_f64.npz"))
    net, _ = build(ast, precision)
    x = weights.content_batch(1, 1080, 2, width=1920).cuda()
    with torch.no_grad():
        y = net(x)
    assert tuple(y.shape) == (1, 3, 1080, 1920)
    sy, sx = 1080 // 32, 1920 // 32
    assert rel(y[:, :, ::sy, ::sx], gold["generated_sub"]) < tol
    assert rel(y.double().sum(dim=3), gold["generated_rowsum"]) < tol
    assert rel(y.double().sum(dim=2), gold["generated_colsum"]) < tol
    assert abs(float(y.double().norm()) - float(gold["generated_norm"])) < tol * float(gold["generated_norm"])


def test_config4_uint8_pipeline_matches_reference_postprocessing(ast):
    """stylize(): uint8 HWC BGR in -> uint8 HWC RGB out, the pre/post-processing of inference.py:110,115-116 fused into
    the end layers.  Bit-exact against applying the reference's numpy post-processing to the float output of the same
    network (clip(0,255).astype('uint8') truncates), and the uint8 input path equals the float input path exactly."""
    net, _ = build(ast, "fast")
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (2, 72, 104, 3), generator=g, dtype=torch.uint8)            # [B,H,W,3] BGR like cv2
    out_u8 = net.stylize(img.cuda()).cpu().numpy()
    with torch.no_grad():
        y = net(img.permute(0, 3, 1, 2).float().cuda()).cpu().numpy()                        # inference.py:110,115
        y_u8in = net(img.permute(0, 3, 1, 2).contiguous().cuda()).cpu().numpy()              # uint8 NCHW input
    assert np.array_equal(y, y_u8in)
    want = y[:, [2, 1, 0]].transpose(0, 2, 3, 1).clip(0, 255).astype("uint8")                # inference.py:116
    assert out_u8.shape == want.shape and out_u8.dtype == np.uint8
    assert np.array_equal(out_u8, want)
    # strict mode goes through the same boundary
    net32, _ = build(ast, "fp32")
    out32 = net32.stylize(img.cuda()).cpu().numpy()
    with torch.no_grad():
        y32 = net32(img.permute(0, 3, 1, 2).float().cuda()).cpu().numpy()
    assert np.array_equal(out32, y32[:, [2, 1, 0]].transpose(0, 2, 3, 1).clip(0, 255).astype("uint8"))


def test_config3_smartaverage_512_fast(ast, golden_dir):
    """'smartaverage' (train_cnn.py:224-244) over 8 paintings at 512^2 in fast mode: reference semantics against the
    reference's own fp32 run; the north-star 'mean of Grams' variant against the oracle port (pinned to the reference's
    gram())."""
    gold = np.load(os.path.join(golden_dir, "smartavg_b1_s512_n8_f32.npz"))
    _, vgg = build(ast, "fast")
    paintings = [weights.style_image(512, 2, i).cuda() for i in range(8)]
    grams = ast.style_grams_smartaverage(vgg, paintings, 1, mode="reference")
    gram_checks({k: v.contiguous() for k, v in grams.items()}, gold, 2e-3)
    alt = ast.style_grams_smartaverage(vgg, paintings, 1, mode="mean_gram")
    vsd = weights.vgg_state_dict(2)
    ref_alt = port.style_grams_smartaverage([p.cpu() for p in paintings], vsd, 1, mode="mean_gram")
    for k in alt:
        assert rel(alt[k], ref_alt[k]) < 2e-3, k


def test_uint8_training_batches_equal_float_batches(ast):
    """uint8 content batches (dataset.py:97-108 data is uint8-derived) give the same step as their float copies."""
    content = weights.content_batch(2, 64, 2)
    res = []
    for dt in (torch.float32, torch.uint8):
        net, vgg = build(ast, "fast")
        style = ast.style_grams_single(vgg, weights.style_image(64, 2).cuda(), 2)
        net.zero_grad()
        c, s, t = ast.perceptual_step(net, vgg, content.to(dt).cuda(), style)
        res.append((float(c), float(s), float(t)))
    np.testing.assert_allclose(np.array(res[0]), np.array(res[1]), rtol=2e-3)    # atomics order only


def test_cycle_gram_bank(ast):
    """'cycle' (train_cnn.py:206-223,316-320): the bank's target t equals the single-image style setup of painting t % P,
    and a trainer step with `style_gram=bank.target(t)` equals a step of a trainer built on that painting."""
    _, vgg = build(ast, "fast")
    paintings = [weights.style_image(64, 2, i).cuda() for i in range(3)]
    bank = ast.StyleGramBank(vgg, paintings)
    assert len(bank) == 3
    for t in (0, 1, 2, 4):
        single = ast.style_grams_single(vgg, paintings[t % 3], 2)
        for k, v in bank.target(t).items():
            assert rel(v, single[k][0]) < 1e-6           # split-K atomics: summation order differs run to run
    content = [weights.content_batch(2, 64, 2, step=i).cuda() for i in range(6)]
    for graph in (False, True):
        neta, _ = build(ast, "fast")
        tr = ast.PerceptualTrainer(neta, vgg, bank.target(0), lr=1e-3, cuda_graph=graph)
        got = [tuple(float(v) for v in tr.step(c, style_gram=bank.target(i))) for i, c in enumerate(content)]
        for i in (1, 5):                       # painting 1 and painting 2: compare with a fresh single-style step
            netb, _ = build(ast, "fast")
            # bring netb to the same parameters: replay the same steps up to i with the same targets
            trb = ast.PerceptualTrainer(netb, vgg, bank.target(0), lr=1e-3)
            ref = [tuple(float(v) for v in trb.step(c, style_gram=ast.style_grams_single(vgg, paintings[j % 3], 2)))
                   for j, c in enumerate(content[:i + 1])]
            np.testing.assert_allclose(np.array(got[i]), np.array(ref[i]), rtol=5e-3)
        tr.close()


@pytest.mark.parametrize("h,w", [(36, 52), (68, 76), (132, 140)])
def test_sizes_that_are_not_multiples_of_16_fast_vs_strict(ast, h, w):
    """Image sizes that are multiples of 4 only (partial tiles in every tensor-core kernel: block-stacked phases and
    interleaved rows, halo tiles, fp16 activations, pooling codes): the fast path agrees with the strict FFMA path of the
    same weights - losses to 1e-2 (north_star), gradients to bf16 accuracy."""
    torch.manual_seed(h * w)
    x = torch.randint(0, 256, (2, 3, h, w), device="cuda").float()
    style = torch.randint(0, 256, (3, h, w), device="cuda").float()
    res = []
    for precision in ("fast", "fp32"):
        net, vgg = build(ast, precision)
        sg = ast.style_grams_single(vgg, style, 2)
        c, s, _ = ast.perceptual_step(net, vgg, x, sg)
        g = torch.cat([p.grad.flatten() for p in net.parameters()])
        res.append((float(c), float(s), g))
    assert abs(res[0][0] - res[1][0]) <= 1e-2 * abs(res[1][0])
    assert abs(res[0][1] - res[1][1]) <= 1e-2 * abs(res[1][1])
    assert rel(res[0][2], res[1][2]) < 5e-2


def test_forward_on_sizes_that_are_not_multiples_of_4(ast):
    """The transform net maps (250, 330) -> (252, 332) like the reference's layers (stride-2 convs round up, the transposed
    convs double); fast and strict mode agree."""
    net_f, _ = build(ast, "fast")
    net_s, _ = build(ast, "fp32")
    x = torch.randint(0, 256, (1, 3, 250, 330), device="cuda").float()
    with torch.no_grad():
        yf, ys = net_f(x), net_s(x)
    assert tuple(yf.shape) == tuple(ys.shape) == (1, 3, 252, 332)
    assert rel(yf, ys) < 5e-2
