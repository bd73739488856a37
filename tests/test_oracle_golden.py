"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU-only."""
import os

import pytest
import torch

from oracle import flamingo_oracle as O
from tests.golden.make_golden import digest

DT = torch.float64


def _check_grads(named_grads, packed, rtol=1e-5):
    assert set(named_grads) == set(packed)
    for n, g in named_grads.items():
        ref = packed[n]
        if "full" in ref:
            torch.testing.assert_close(g.float(), ref["full"], rtol=rtol, atol=1e-6, msg=lambda m: f"{n}: {m}")
        else:
            d = digest(n, g)
            torch.testing.assert_close(d, ref["digest"], rtol=1e-6, atol=1e-6, msg=lambda m: f"{n}: {m}")


@pytest.mark.parametrize("name", ["res_img", "res_vid", "res_relu", "res_h4"])
def test_resampler_matches_reference(golden_dir, name):
    fx = torch.load(os.path.join(golden_dir, f"{name}.pt"))
    c = fx["case"]
    heads, ff_mult = c.get("heads", 8), c.get("ff_mult", 4)
    p = O.seeded_params(O.resampler_param_shapes(c["dim"], c["depth"], heads=heads, ff_mult=ff_mult), c["seed"], dtype=DT)
    assert sum(t.numel() for t in p.values()) == fx["n_params"]
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    x = fx["x"].to(DT).requires_grad_(True)
    out = O.perceiver_resampler(x, p, c["depth"], heads=heads, act=c["act"])
    torch.testing.assert_close(out.float(), fx["out"], rtol=1e-5, atol=1e-5)
    (out * fx["cot"].to(DT)).sum().backward()
    torch.testing.assert_close(x.grad.float(), fx["dx"], rtol=1e-4, atol=1e-5)
    _check_grads({k: v.grad for k, v in p.items()}, fx["dparams"])


@pytest.mark.parametrize("name", ["xattn_edge", "xattn_sq", "xattn_h2"])
def test_xattn_matches_reference(golden_dir, name):
    fx = torch.load(os.path.join(golden_dir, f"{name}.pt"))
    c = fx["case"]
    heads, ff_mult = c.get("heads", 8), c.get("ff_mult", 4)
    p = O.seeded_params(O.xattn_param_shapes(c["dim"], c["dim_visual"], heads=heads, ff_mult=ff_mult), c["seed"], dtype=DT)
    assert sum(t.numel() for t in p.values()) == fx["n_params"]
    p = {k: v.requires_grad_(True) for k, v in p.items()}
    y = fx["y"].to(DT).requires_grad_(True)
    vis = fx["vis"].to(DT).requires_grad_(True)
    out, (k, v) = O.gated_xattn_block(y, vis, fx["media_locations"], p, heads=heads, act=c["act"], output_kv=True)
    torch.testing.assert_close(out.float(), fx["out"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(k.float(), fx["k"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(v.float(), fx["v"], rtol=1e-5, atol=1e-5)
    (out * fx["cot"].to(DT)).sum().backward()
    torch.testing.assert_close(y.grad.float(), fx["dy"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(vis.grad.float(), fx["dvis"], rtol=1e-4, atol=1e-5)
    _check_grads({k_: v_.grad for k_, v_ in p.items()}, fx["dparams"])
    # cached path: last three tokens with the stored keys/values
    with torch.no_grad():
        oc, _ = O.gated_xattn_block(y[:, -3:].detach(), None, fx["media_locations"], p, heads=heads, act=c["act"],
                                    previous_kv=(k.detach(), v.detach()))
    torch.testing.assert_close(oc.float(), fx["out_cached_last3"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(oc.float(), fx["out"][:, -3:], rtol=1e-5, atol=1e-5)


def test_semantic_kats():
    """SURVEY §8a probed facts: gate 0 => identity; tt==0 => zero attention; tt>n_media => uniform; T>4 raises."""
    p = O.seeded_params(O.xattn_param_shapes(64, 64), 5, alpha=0.0)
    y = torch.randn(2, 6, 64)
    vis = torch.randn(2, 1, 64, 64)
    ml = torch.tensor([[0, 0, 1, 0, 0, 0], [1, 0, 0, 1, 0, 0]])
    out, _ = O.gated_xattn_block(y, vis, ml, p)
    assert torch.equal(out, y)
    p = O.seeded_params(O.xattn_param_shapes(64, 64), 5)
    att, _ = O.masked_cross_attention(y, ml, vis, p, "attn.")
    assert torch.equal(att[0, :2], torch.zeros_like(att[0, :2]))            # before the first image
    kv = vis.reshape(2, 64, 64) @ p["attn.to_kv.weight"].T
    uniform = kv[1, :, 512:].mean(0) @ p["attn.to_out.weight"].T            # 2 markers, 1 image
    torch.testing.assert_close(att[1, 3], uniform, rtol=1e-4, atol=1e-5)
    rp = O.seeded_params(O.resampler_param_shapes(64, 1), 3)
    with pytest.raises(RuntimeError):
        O.perceiver_resampler(torch.randn(1, 5, 3, 64), rp, 1)


def test_param_count_kats():
    """examples/model_stats.ipynb:1583-1584,1605,106-111."""
    n_res = sum(torch.Size(s).numel() for s in O.resampler_param_shapes(1024, 6).values())
    n_x = sum(torch.Size(s).numel() for s in O.xattn_param_shapes(768, 1024).values())
    assert n_res == 63_023_104
    assert n_res + 12 * n_x + 50_273 * 768 == 180_312_856
    assert len(O.resampler_param_shapes(1024, 6)) + 1 + 12 * len(O.xattn_param_shapes(768, 1024)) == 209
