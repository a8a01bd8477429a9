"""GPU parity of the tensor-core pair kernels (csrc/nais_pairs_tc.cu, csrc/nais_pairs_tc_bwd.cu — what `forward` / `backward` run by
default, NaisParams::pairs_precision = NAIS_PAIRS_AUTO): the same C-ABI entry points, the same oracle and tolerance as the FP32
kernels' tests, plus closeness to the FP32 kernels themselves (model.pairs_precision = "fp32")."""
import os

import numpy as np
import pytest
import torch

import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import synthetic

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(variant, N, D, hid, B, H, seed, style="trained"):
    rng = np.random.default_rng(seed)
    coords, region, R = synthetic.make_catalog(N, seed=seed)
    sd = orc.init_state(variant, N, D, hid, R, 1, seed=seed + 1, style=style)
    hist = np.stack([rng.choice(N, H, replace=False) for _ in range(B)]).astype(np.int64)
    tgt = rng.integers(0, N, B).astype(np.int64)
    if H > 1:
        tgt[::3] = hist[::3, H // 2]  # live mask, like a training positive
    aux = orc.latlon_abs_diff(coords, tgt, hist) if orc.VARIANTS[variant]["dist"] == "latlon" else None
    return sd, hist, tgt, region, aux


def _both(variant, sd, beta, hist, tgt, region, aux):
    m = util.make_model(variant, sd, beta)
    args = (_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), None if aux is None else _dev(aux))
    with torch.no_grad():
        m.pairs_precision = "tc"  # NAIS_PAIRS_TC: an unsupported shape is an error, not a silent FP32 run
        s_tc = util.call(m, variant, *args).cpu().numpy()
        m.pairs_precision = "auto"
        assert np.array_equal(util.call(m, variant, *args).cpu().numpy(), s_tc), "auto must pick the tensor-core kernel here"
        m.pairs_precision = "fp32"
        s_fp = util.call(m, variant, *args).cpu().numpy()
    ref, scale = orc.attention_network_with_scale(sd, variant, beta, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                  torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                  None if aux is None else torch.from_numpy(aux), dtype=torch.float64)
    return s_tc, s_fp, ref.numpy(), scale.numpy()


@pytest.mark.parametrize("H", [1, 3, 13, 100, 128, 129, 300])
def test_pairs_tc_history_lengths(H):
    sd, hist, tgt, region, aux = _case("region_distance", 900, 64, 64, 37, H, seed=H)
    s_tc, s_fp, ref, scale = _both("region_distance", sd, 0.5, hist, tgt, region, aux)
    assert not np.array_equal(s_tc, s_fp) or H == 1, "the tensor-core path did not run (results bit-identical to FP32)"
    assert util.cond_err(s_tc, ref, scale) < util.TOL
    assert util.cond_err(s_tc, ref, scale) <= max(4 * util.cond_err(s_fp, ref, scale), 5e-6)  # fp32-grade, not just < 1e-4


@pytest.mark.parametrize("variant,D,hid", [("region_distance", 32, 32), ("region_distance", 64, 128), ("region_distance", 16, 48),
                                           ("basic", 64, 64), ("region", 48, 64), ("distance", 64, 64)])
def test_pairs_tc_variants_and_shapes(variant, D, hid):
    sd, hist, tgt, region, aux = _case(variant, 500, D, hid, 300, 21, seed=D + hid)
    s_tc, s_fp, ref, scale = _both(variant, sd, 0.7, hist, tgt, region, aux)
    assert util.cond_err(s_tc, ref, scale) < util.TOL
    assert util.cond_err(s_tc, ref, scale) <= max(4 * util.cond_err(s_fp, ref, scale), 5e-6)


@pytest.mark.parametrize("scale_e,scale_w", [(1e-4, 1.0), (30.0, 0.01), (1.0, 50.0)])
def test_pairs_tc_row_scaling_extremes(scale_e, scale_w):
    """Per-row power-of-two scaling of X and the global one of W: tiny / huge embeddings and weights keep fp32-grade products."""
    sd, hist, tgt, region, aux = _case("region_distance", 400, 64, 64, 64, 50, seed=9)
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if k.startswith("embed_"):
            sd[k] *= scale_e
    sd["attn_layer1.weight"][:, :64] *= scale_w
    s_tc, s_fp, ref, scale = _both("region_distance", sd, 0.5, hist, tgt, region, aux)
    ok = np.isfinite(ref)
    assert ok.any()
    assert util.cond_err(s_tc[ok], ref[ok], scale[ok]) < util.TOL
    assert util.cond_err(s_tc[ok], ref[ok], scale[ok]) <= max(4 * util.cond_err(s_fp[ok], ref[ok], scale[ok]), 5e-6)


@pytest.mark.parametrize("bwd", ["tc", "fp32_backward"])
def test_pairs_tc_c3_shape_forward_and_gradients(bwd):
    """C3-shaped batch (own history per row, H = 128, D = hid = 64): forward vs the float64 oracle on every row, and the
    gradients vs the oracle's autograd — tensor-core forward + tensor-core backward (the default pairing), and the FP32
    backward fed by the tensor-core forward's saved row sums."""
    B, H, N, D, hid, beta = 512, 128, 3000, 64, 64, 0.5
    sd, hist, tgt, region, aux = _case("region_distance", N, D, hid, B, H, seed=21)
    rng = np.random.default_rng(5)
    dscore = rng.normal(size=B)
    m = util.make_model("region_distance", sd, beta)
    m.pairs_precision = "tc" if bwd == "tc" else ("tc", "fp32")  # (forward, backward)
    s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux))
    (s * _dev(dscore).float()).sum().backward()
    ref_s, ref = orc.grads(sd, "region_distance", beta, torch.from_numpy(hist), torch.from_numpy(tgt), torch.from_numpy(region[hist]),
                           torch.from_numpy(region[tgt]), torch.from_numpy(aux), torch.from_numpy(dscore))
    _, scale = orc.attention_network_with_scale(sd, "region_distance", beta, torch.from_numpy(hist), torch.from_numpy(tgt),
                                                torch.from_numpy(region[hist]), torch.from_numpy(region[tgt]),
                                                torch.from_numpy(aux), dtype=torch.float64)
    assert util.cond_err(s.detach().cpu().numpy(), ref_s.numpy(), scale.numpy()) < util.TOL
    ref = {k: v.numpy() for k, v in ref.items()}
    ref.setdefault("embed_distance.weight", np.zeros((1, D)))
    for name, p in m.named_parameters():
        r = np.asarray(ref[name], dtype=np.float64)
        g = np.zeros_like(r) if p.grad is None else p.grad.detach().cpu().double().numpy()
        assert np.abs(g - r).max() <= 2e-4 * np.abs(r).max() + 1e-12, name


def _grads_gpu(sd, beta, hist, tgt, region, aux, dscore, precision):
    m = util.make_model("region_distance", sd, beta)
    m.pairs_precision = precision
    s = m.attention_network(_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux))
    (s * _dev(dscore).float()).sum().backward()
    torch.cuda.synchronize()
    return {n: (None if p.grad is None else p.grad.detach().cpu().double().numpy()) for n, p in m.named_parameters()}


@pytest.mark.parametrize("B,H,D", [(256, 128, 64), (150, 21, 64), (40, 300, 64), (100, 50, 32)])
def test_pairs_tc_backward_matches_oracle(B, H, D):
    """Tensor-core backward (csrc/nais_pairs_tc_bwd.cu: bf16 two-term splits, dW accumulated in TMEM through
    MN-major reads of the operand images) against the oracle's float64 autograd, with the gradient bar of the FP32 backward's
    tests (per tensor 2e-4 of its maximum); and it must not be the FP32 backward in disguise."""
    sd, hist, tgt, region, aux = _case("region_distance", 3000, D, 64, B, H, seed=B + H)
    dscore = np.random.default_rng(7).normal(size=B)
    g_tc = _grads_gpu(sd, 0.5, hist, tgt, region, aux, dscore, "tc")
    g_fp = _grads_gpu(sd, 0.5, hist, tgt, region, aux, dscore, "fp32")
    _, ref = orc.grads(sd, "region_distance", 0.5, torch.from_numpy(hist), torch.from_numpy(tgt), torch.from_numpy(region[hist]),
                       torch.from_numpy(region[tgt]), torch.from_numpy(aux), torch.from_numpy(dscore))
    ref = {k: v.numpy() for k, v in ref.items()}
    ref.setdefault("embed_distance.weight", np.zeros((1, D)))
    assert not np.array_equal(g_tc["attn_layer1.weight"], g_fp["attn_layer1.weight"]), "the tensor-core backward did not run"
    worst = {}
    for name, r in ref.items():
        r = np.asarray(r, dtype=np.float64)
        g = np.zeros_like(r) if g_tc.get(name) is None else g_tc[name]
        worst[name] = float(np.abs(g - r).max() / max(np.abs(r).max(), 1e-300))
    print("tc backward, max err / max|ref| per tensor:", {k: f"{v:.1e}" for k, v in worst.items()})
    for name, v in worst.items():
        if np.abs(ref[name]).max() > 0:
            assert v <= 2e-4, (name, v)


def test_pairs_tc_backward_timing(capsys):
    """Not a pass/fail timing: prints the C3-shaped backward (4096 rows x H = 128) in both modes for the round log."""
    sd, hist, tgt, region, aux = _case("region_distance", 40000, 64, 64, 4096, 128, seed=3)
    m = util.make_model("region_distance", sd, 0.5)
    args = (_dev(hist), _dev(tgt), _dev(region[hist]), _dev(region[tgt]), _dev(aux))
    out = {}
    for mode in ("fp32", "tc"):
        m.pairs_precision = mode
        ts = []
        for it in range(4):
            m.zero_grad(set_to_none=True)
            s = m.attention_network(*args).sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            s.backward()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[mode] = min(ts[1:])
    with capsys.disabled():
        print(f"\nC3-shaped backward incl. reduces: fp32 {out['fp32']:.3f} ms, tensor-core {out['tc']:.3f} ms")
