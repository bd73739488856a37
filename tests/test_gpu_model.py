"""Model-level parity ON THE GPU: FlamingoModel with the CUDA PerceiverResampler and
GatedCrossAttentionBlocks spliced into the stock HF language model, against tensors recorded from the UNMODIFIED reference
FlamingoModel (tests/golden/make_golden_model.py; modeling_flamingo.py:183-306, gated_cross_attention.py:231-252).

What is compared: logits, loss, EVERY trainable gradient (resampler, gated xattn blocks, token embedding), the cached prefix
forward and one cached decode step (S=1 through previous_kv).  The frozen LM / CLIP run in fp32 on the GPU (stock PyTorch), so
the differences measured here are the hot-path kernels' (bf16 tensor-core operands, fp32 accumulation).

Tolerances (stated, element-wise): logits |got - ref| <= ATOL + RTOL * |ref| with RTOL = 1e-2, ATOL = 2e-2 * rms(ref logits)
plus a whole-tensor relative L2 bound of 5e-3; gradients relative L2 <= 6e-2 per tensor (same bound as the module tests).
The measured values are printed and summarised in profiles/r02_parity.md.
"""
import os

import pytest
import torch

from flamingo_mini_b200.configuration_flamingo import FlamingoConfig
from flamingo_mini_b200.modeling_flamingo import FlamingoModel
from tests._gpu_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

LOGIT_RTOL, LOGIT_ATOL_RMS, LOGIT_L2 = 1e-2, 2e-2, 5e-3
GRAD_L2 = 6e-2


def _build(fx, which):
    if which == "opt":
        cfg = FlamingoConfig(lm="facebook/opt-125m", dim=64, dim_visual=64, xattn_every=1, resampler_depth=1,
                             lm_config=fx["opt_cfg"], clip_config=fx["clip_cfg"])
    else:
        cfg = FlamingoConfig(lm="gpt2", dim=64, dim_visual=64, xattn_every=2, resampler_depth=1, xattn_act="sqrelu",
                             lm_config=fx["gpt2_cfg"], clip_config=fx["clip_cfg"])
    model = FlamingoModel(cfg)
    res = model.load_state_dict(fx["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return model.to(DEV).eval()


def _logits_close(got, ref, what):
    ref = ref.to(got.device).float()
    got = got.float()
    rms = ref.square().mean().sqrt().item()
    l2 = rel_err(got, ref)
    worst = ((got - ref).abs() - LOGIT_RTOL * ref.abs()).max().item()
    print(f"[parity] {what}: rel L2 {l2:.3e}, max abs err {(got - ref).abs().max().item():.3e}, rms(ref) {rms:.3e}")
    assert l2 <= LOGIT_L2, f"{what}: relative L2 error {l2:.3e} > {LOGIT_L2}"
    torch.testing.assert_close(got, ref, rtol=LOGIT_RTOL, atol=LOGIT_ATOL_RMS * rms, msg=lambda m: f"{what}: {m} (slack {worst:.3e})")


@pytest.mark.parametrize("which", ["opt", "gpt2"])
def test_flamingo_model_matches_reference_on_gpu(golden_dir, which):
    fx = torch.load(os.path.join(golden_dir, f"model_{which}_tiny.pt"))
    model = _build(fx, which)
    from flamingo_mini_b200 import _lib
    n0 = _lib.load().fm_launch_count()
    ids, ml, pix = fx["input_ids"].to(DEV), fx["media_locations"].to(DEV), fx["pixel_values"].to(DEV)
    out = model(input_ids=ids, media_locations=ml, pixel_values=pix, labels=ids, attention_mask=torch.ones_like(ids))
    assert _lib.load().fm_launch_count() > n0, "the CUDA library launched nothing: the hot path did not run on the kernels"
    _logits_close(out.logits, fx["logits"], f"{which} logits")
    torch.testing.assert_close(out.loss.float().cpu(), fx["loss"], rtol=2e-3, atol=2e-3)
    out.loss.backward()
    got = {n: p.grad for n, p in model.named_parameters() if p.requires_grad}
    assert set(got) == set(fx["grads"]), sorted(set(got) ^ set(fx["grads"]))
    worst = ("", 0.0)
    for n, ref in fx["grads"].items():
        assert got[n] is not None, f"no gradient for {n}"
        ref = ref.to(DEV)
        if ref.norm().item() < 1e-7:
            assert got[n].float().norm().item() < 1e-4, n
            continue
        e = rel_err(got[n], ref)
        if e > worst[1]:
            worst = (n, e)
        # the two gate gradients are scalar sums of signed products (random-walk error, see test_gpu_modules): wider bound
        tol = 0.15 if ".alpha_" in n else GRAD_L2
        assert e <= tol, f"grad {n}: rel L2 err {e:.3e} > {tol}"
    print(f"[parity] {which} worst gradient: {worst[0]} rel L2 {worst[1]:.3e} over {len(fx['grads'])} tensors")


@pytest.mark.parametrize("which", ["opt", "gpt2"])
def test_cached_prefix_and_decode_step_on_gpu(golden_dir, which):
    """prefix forward with use_cache, then ONE decode step (S = 1) through (xattn_past, lm_past): gated_cross_attention.py:88-104,
    modeling_flamingo.py:238-239,282-285,303.  The 1-token step runs the decode-shape xattn path."""
    fx = torch.load(os.path.join(golden_dir, f"model_{which}_tiny.pt"))
    model = _build(fx, which)
    ids, ml, pix = fx["input_ids"].to(DEV), fx["media_locations"].to(DEV), fx["pixel_values"].to(DEV)
    with torch.no_grad():
        first = model(input_ids=ids[:, :8], media_locations=ml[:, :8], pixel_values=pix, use_cache=True,
                      attention_mask=torch.ones_like(ids[:, :8]))
        _logits_close(first.logits, fx["logits_prefix"], f"{which} prefix logits")
        k0 = first.past_key_values[0][0][0]
        if "cache_k_shape" in fx:
            assert tuple(k0.shape) == fx["cache_k_shape"]
        step = model(input_ids=ids[:, 8:9], media_locations=ml[:, :9], past_key_values=first.past_key_values, use_cache=True,
                     attention_mask=torch.ones_like(ids[:, :9]))
        _logits_close(step.logits, fx["logits_step"], f"{which} cached decode-step logits")
        # the cached step must also agree with the un-cached forward over 9 tokens, position 8
        full = model(input_ids=ids[:, :9], media_locations=ml[:, :9], pixel_values=pix, attention_mask=torch.ones_like(ids[:, :9]))
        _logits_close(step.logits[:, 0], full.logits[:, 8].cpu(), f"{which} cached step vs uncached")


def test_generate_on_gpu_cached_equals_uncached(golden_dir):
    """generate() (what generate_captions calls, modeling_flamingo.py:550-605) with the CUDA modules: cached greedy decoding
    picks the tokens an un-cached full forward picks wherever the un-cached top-2 margin exceeds the bf16 noise floor."""
    fx = torch.load(os.path.join(golden_dir, "model_opt_tiny.pt"))
    model = _build(fx, "opt")
    n0, n1 = 5, 12
    ids, ml, pix = fx["input_ids"][:, :n0].to(DEV), fx["media_locations"][:, :n0].to(DEV), fx["pixel_values"].to(DEV)
    with torch.no_grad():
        greedy = model.generate(inputs=ids, media_locations=ml, attention_mask=torch.ones_like(ids), pixel_values=pix, use_cache=True,
                                max_length=n1, do_sample=False, pad_token_id=0, eos_token_id=None)
        assert greedy.shape == (2, n1) and torch.equal(greedy[:, :n0], ids)
        cml = torch.cat([ml, torch.zeros(2, n1 - n0, dtype=ml.dtype, device=DEV)], 1)
        lg = model(input_ids=greedy, media_locations=cml, attention_mask=torch.ones_like(greedy), pixel_values=pix).logits.float()
        for t in range(n0, n1):
            top2 = lg[:, t - 1].topk(2).values
            pick = lg[:, t - 1].argmax(-1)
            sure = (top2[:, 0] - top2[:, 1]) > 5e-2 * lg[:, t - 1].abs().max()
            assert torch.equal(greedy[sure, t], pick[sure]), f"position {t}: cached greedy token differs from the un-cached argmax"
