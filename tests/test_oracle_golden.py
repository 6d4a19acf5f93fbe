"""Pins the CPU oracle (oracle/port.py) to outputs of the real reference classes
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import port, weights


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _gram_summary(g):
    g = g.detach().double()
    return {"block": g[:, :32, :32].numpy(), "diag": torch.diagonal(g, dim1=1, dim2=2).numpy(),
            "fro": g.flatten(1).norm(dim=1).numpy(), "sum": g.flatten(1).sum(dim=1).numpy()}


def _run_oracle_step(batch, size, dtype, grads=True):
    tsd = port.make_leaf(weights.transfer_state_dict(2), dtype)
    vsd = {k: v.to(dtype) for k, v in weights.vgg_state_dict(2).items()}
    content = weights.content_batch(batch, size, 2).to(dtype)
    style = port.style_grams_single(weights.style_image(size, 2).to(dtype), vsd, batch)
    return port.training_step(tsd, vsd, content, style, want_grads=grads), style


@pytest.mark.parametrize("name,batch,size,dtype,rtol", [
    ("step_b2_s32_f64", 2, 32, torch.float64, 1e-10),
    ("step_b2_s64_f64", 2, 64, torch.float64, 1e-10),
])
def test_oracle_step_matches_reference_fp64(golden_dir, name, batch, size, dtype, rtol):
    gold = _load(golden_dir, name)
    out, style = _run_oracle_step(batch, size, dtype)
    got = np.array([float(out["content"]), float(out["style"]), float(out["total"])])
    np.testing.assert_allclose(got, gold["losses"], rtol=rtol)
    for k, g in out["grams"].items():
        for n, v in _gram_summary(g).items():
            np.testing.assert_allclose(v, gold[f"gram/{k}/{n}"], rtol=1e-9, atol=1e-12)
    for k, g in style.items():
        for n, v in _gram_summary(g).items():
            np.testing.assert_allclose(v, gold[f"style_gram/{k}/{n}"], rtol=1e-9, atol=1e-12)
    sub = out["generated"].double()[:, :, ::max(1, size // 32), ::max(1, size // 32)].numpy()
    np.testing.assert_allclose(sub, gold["generated_sub"], rtol=1e-9, atol=1e-10)
    # gradients: fp64 against fp64 is well conditioned (SURVEY 8c noise-floor table is fp32-vs-fp64)
    for key, g in out["grads"].items():
        ref_norm = float(gold["grad_norm/" + key])
        # conv biases feeding an InstanceNorm are mathematically dead (|g| ~ 1e-14 noise)
        np.testing.assert_allclose(float(g.double().norm()), ref_norm, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(g.double().flatten()[:16].numpy(), gold["grad_head/" + key],
                                   rtol=1e-6, atol=1e-10)


def test_oracle_step_config1_fp32(golden_dir):
    """BASELINE config 1 (B=4, 256^2) in fp32 against the reference's fp32 AND fp64 runs."""
    g32 = _load(golden_dir, "step_b4_s256_f32")
    g64 = _load(golden_dir, "step_b4_s256_f64")
    out, _ = _run_oracle_step(4, 256, torch.float32, grads=False)
    got = np.array([float(out["content"]), float(out["style"]), float(out["total"])])
    np.testing.assert_allclose(got, g32["losses"], rtol=2e-6)
    np.testing.assert_allclose(got, g64["losses"], rtol=1e-5)      # north_star strict tolerance
    for k, g in out["grams"].items():
        s = _gram_summary(g)
        ref = g64[f"gram/{k}/block"]
        err = np.linalg.norm(s["block"] - ref) / np.linalg.norm(ref)
        assert err < 1e-5, (k, err)
        np.testing.assert_allclose(s["fro"], g64[f"gram/{k}/fro"], rtol=1e-5)


def test_oracle_smartaverage(golden_dir):
    gold = _load(golden_dir, "smartavg_b2_s64_n5_f64")
    vsd = {k: v.double() for k, v in weights.vgg_state_dict(2).items()}
    paintings = [weights.style_image(64, 2, i).double() for i in range(5)]
    grams = port.style_grams_smartaverage(paintings, vsd, 2, mode="reference")
    for k, g in grams.items():
        for n, v in _gram_summary(g).items():
            np.testing.assert_allclose(v, gold[f"gram/{k}/{n}"], rtol=1e-9, atol=1e-12)
    # the north-star 'mean of Grams' variant is a different quantity
    alt = port.style_grams_smartaverage(paintings, vsd, 2, mode="mean_gram")
    assert not np.allclose(alt["relu1_2"].numpy(), grams["relu1_2"].numpy(), rtol=1e-3)


def test_state_dict_layout():
    sd = weights.transfer_state_dict(2)
    assert len(sd) == 70
    assert sum(v.numel() for v in sd.values()) == 1712771
    assert sd["DeconvBlock.2.conv_transpose.weight"].shape == (128, 64, 3, 3)
    assert list(sd.keys()) == port.transfer_param_keys()
