"""The library (csrc/) executed on the host emulator: the bodies of the -m gpu tests
are reused unchanged with DEV = "cpu" and the emulated libflamingo_b200_emu.so swapped in (tests/_emu_util.py).
What a pass means: kernel logic — tile schedules, mbarrier protocols, TMA boxes, UMMA descriptors, TMEM addressing,
epilogue indexing, reductions — is right under the functional model of tests/cpu_harness/tc_emu.h.  What it does not
mean: anything about speed, or about hardware behaviour outside that model."""
import os

import pytest
import torch

from tests import _emu_util

pytestmark = pytest.mark.skipif(not _emu_util.available(), reason="needs g++ and the CUDA headers")


@pytest.fixture()
def emu(monkeypatch):
    import tests.test_gpu_gemm as G
    import tests.test_gpu_modules as M
    import tests.test_gpu_ops as P
    for mod in (G, M, P):
        monkeypatch.setattr(mod, "DEV", "cpu")
        if hasattr(mod, "stream"):
            monkeypatch.setattr(mod, "stream", lambda: None)
    with _emu_util.swapped_in() as lib:
        yield lib


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 0), (256, 256, 256, 64), (256, 256, 256, 192), (1000, 520, 200, 0),
                                      (1000, 520, 200, 128), (1000, 520, 200, 256), (136, 8, 72, 0)])
def test_gemm_store_f32(emu, a_mn, b_mn, M, N, K, bn):
    import tests.test_gpu_gemm as G
    G.test_gemm_store_f32(a_mn, b_mn, M, N, K, bn)


def test_gemm_epilogues(emu):
    import tests.test_gpu_gemm as G
    G.test_gemm_store_bf16_scale_bias_gate()
    for act in (0, 1, 2):
        G.test_gemm_act_epilogue(act)
    for aux_f32, out_f32 in [(0, 1), (1, 1), (1, 0), (0, 0)]:
        G.test_gemm_resid_epilogue(aux_f32, out_f32)
    G.test_gemm_dact_epilogue()


@pytest.mark.parametrize("M,N,K", [(768, 512, 2048), (130 * 8, 200, 1000)])
@pytest.mark.parametrize("splits", [0, 2, 3])
def test_gemm_serial_split_k(emu, M, N, K, splits):
    """CTAs of one launch wait for each other through global flags: the emulator keeps all of them resident."""
    import tests.test_gpu_gemm as G
    G.test_gemm_dw_split_k(M, N, K, splits)


@pytest.mark.parametrize("M,N,K,splits", [(392, 32, 48, 2), (480, 408, 240, 3), (256, 64, 64, 4)])
def test_gemm_split_k_without_empty_ranges(emu, M, N, K, splits):
    import tests.test_gpu_gemm as G
    G.test_gemm_split_k_without_empty_ranges(M, N, K, splits)


SMALL_GROUPS = {      # the shapes the modules group (q + kv, dyn + dvis, dWout + dWq + dWkv) at emulator-friendly sizes
    "fwd_small": [(0, 0, 520, 512, 256), (0, 0, 264, 1024, 256)],
    "dx_small": [(0, 1, 520, 256, 512), (0, 1, 264, 256, 1024)],
    "dw_small": [(1, 1, 768, 512, 512), (1, 1, 512, 768, 512), (1, 1, 1024, 768, 256)],
}


@pytest.mark.parametrize("name", ["fwd_small", "dx_small", "dw_small", "ragged", "single"])
@pytest.mark.parametrize("bn,grouped", [(0, 1), (64, 1), (256, 1), (0, 0)])
def test_gemm_group(emu, monkeypatch, name, bn, grouped):
    """One persistent launch over up to four problems; with a forced tile width the results must equal single launches bit for bit."""
    import tests.test_gpu_gemm as G
    for k, v in SMALL_GROUPS.items():
        monkeypatch.setitem(G.GROUPS, k, v)
    G.test_gemm_group(name, bn, grouped)


def test_gemm_store_reduction(emu, monkeypatch):
    import tests.test_gpu_gemm as G
    G.test_gemm_store_reduction()


# ------------------------------------------------------------------------------------------------ CUDA-core kernels through the ABI
def test_layernorm_text_time_cast_and_loss(emu):
    import tests.test_gpu_ops as P
    for rows, D in [(37, 64), (40, 768)]:
        for x_f32 in (0, 1):
            P.test_layernorm_fwd_bwd(rows, D, x_f32)
    P.test_text_time_and_cast()
    for rows, vocab, ld in [(7, 1000, 1000), (33, 515, 520), (5, 8, 8)]:
        P.test_cross_entropy_vs_torch(rows, vocab, ld)


# ------------------------------------------------------------------------------------------------ the two modules, fwd + bwd
@pytest.mark.parametrize("name", ["xattn_edge", "xattn_sq", "xattn_h2"])
def test_xattn_block_against_the_reference_golden(emu, golden_dir, name):
    """GatedCrossAttentionBlock forward / backward / cached decoding on the emulator vs vectors produced by the unmodified
    reference (tests/golden): text before any image, more <image> tags than images, sqrelu, 2 heads with ff_mult 2."""
    import tests.test_gpu_modules as M
    M.test_xattn_golden(golden_dir, name, torch.float32)


@pytest.mark.parametrize("name", ["res_img", "res_vid", "res_relu", "res_h4"])
def test_resampler_against_the_reference_golden(emu, golden_dir, name):
    import tests.test_gpu_modules as M
    M.test_resampler_golden(golden_dir, name, torch.bfloat16)


def test_modules_seeded_vs_oracle(emu):
    import tests.test_gpu_modules as M
    M.test_xattn_identity_at_zero_gate()
    M.test_xattn_seeded_vs_oracle(2, 150, 2, 128, 64)
    M.test_resampler_seeded_vs_oracle(2, 2, 33, 128, 1)
    M.test_resampler_rejects_too_many_frames()
    M.test_parameters_cast_before_the_first_forward()


SLOW = bool(os.environ.get("FM_EMU_SLOW"))      # the default CPU suite keeps one representative of each sweep (a few minutes in total)


@pytest.mark.parametrize("heads", [1, 4, 12] if SLOW else [12])
def test_other_head_counts(emu, heads):
    import tests.test_gpu_modules as M
    M.test_modules_with_other_head_counts(heads)


def test_scheduling_switches_do_not_change_results(emu):
    """fm_set_option: grouped launches off, epilogue L2 prefetch off, d(alpha_ffw) from DACT instead of the dW2 epilogue,
    LayerNorm folds on the main stream, side stream off — every combination must still match the oracle."""
    import tests.test_gpu_modules as M
    from tests._gpu_util import set_option
    sweeps = [dict(gemm_group=0, alpha_from_dw2=0, dattn_from_gemm=0, attn_tmem_compact=0), dict(side_stream=0, epi_prefetch=0, ln_reduce_side=0, pdl=1)]
    for opts in sweeps:
        try:
            for k, v in opts.items():
                assert set_option(k, v)
            M.test_xattn_seeded_vs_oracle(2, 150, 2, 128, 64)
            M.test_resampler_seeded_vs_oracle(2, 1, 20, 128, 1)
        finally:
            for k in opts:
                set_option(k, M.OPTION_DEFAULTS[k])


def test_deferred_side_join_bookkeeping(emu):
    """Host logic of defer_join (streams are no-ops on the emulator): parked buffers, join, accumulation fallback, same gradients."""
    import tests.test_gpu_modules as M
    M.test_deferred_side_join()


def test_standalone_forwards(emu):
    """FeedForward / MaskedCrossAttention / PerceiverAttentionLayer called on their own (standalone.py: primitives + the
    attention cores the C ABI exports), incl. cached decoding and the two degenerate masking rows."""
    import tests.test_gpu_modules as M
    M.test_feed_forward_standalone("gelu", torch.bfloat16)
    if SLOW:
        M.test_feed_forward_standalone("sqrelu", torch.float32)
    for heads in ((8, 3) if SLOW else (3,)):
        M.test_masked_cross_attention_standalone(heads)
        M.test_perceiver_attention_standalone(heads)


@pytest.mark.parametrize("heads", [8, 2] if SLOW else [2])
def test_standalone_attention_modules_with_gradients(emu, heads):
    import tests.test_gpu_modules as M
    M.test_standalone_attention_modules_with_gradients(heads)


def test_whole_model_through_the_emulated_library(emu, golden_dir):
    """The reference FlamingoModel fixture (OPT branch, tests/golden/make_golden_model.py) with OUR fused modules running on the
    emulated library: conditioning, the LM splice, loss, backward into the flat arenas and one cached decoding step.
    The modules compute in bf16 (fp32 residual stream in, fp32 out), hence the tolerances."""
    import os
    from tests.test_model_golden_cpu import _build
    fx = torch.load(os.path.join(golden_dir, "model_opt_tiny.pt"))
    model = _build(fx).eval()
    ids, ml, pix = fx["input_ids"], fx["media_locations"], fx["pixel_values"]
    out = model(input_ids=ids, media_locations=ml, pixel_values=pix, labels=ids, attention_mask=torch.ones_like(ids))
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()      # noqa: E731
    assert rel(out.logits, fx["logits"]) < 2e-2
    assert abs(out.loss.item() - fx["loss"].item()) < 2e-2 * abs(fx["loss"].item())
    out.loss.backward()
    blk = model.flamingo.lm.decoder.layers[0].xattn_block
    assert rel(model.flamingo.resampler.latents.grad, fx["grad_latents"]) < 8e-2
    assert blk.alpha_attn.grad is not None and torch.isfinite(blk.alpha_attn.grad).all()
    with torch.no_grad():
        first = model(input_ids=ids[:, :8], media_locations=ml[:, :8], pixel_values=pix, use_cache=True,
                      attention_mask=torch.ones_like(ids[:, :8]))
        assert tuple(first.past_key_values[0][0][0].shape) == fx["cache_k_shape"]
        assert rel(first.logits, fx["logits_prefix"]) < 2e-2
        step = model(input_ids=ids[:, 8:9], media_locations=ml[:, :9], past_key_values=first.past_key_values, use_cache=True,
                     attention_mask=torch.ones_like(ids[:, :9]))
        full = model(input_ids=ids[:, :9], media_locations=ml[:, :9], pixel_values=pix, attention_mask=torch.ones_like(ids[:, :9]))
        assert rel(step.logits[:, -1], full.logits[:, -1]) < 2e-2


@pytest.mark.skipif(not os.environ.get("FM_EMU_SLOW"), reason="~6 min on the emulator (depth-6 resampler, 5 optimiser steps): set FM_EMU_SLOW=1")
def test_training_steps_through_the_emulated_library(emu):
    """training.train (fwd, bwd into the flat arenas, ArenaAdamW) on the tiny workload: the body of the GPU test, on CPU tensors."""
    import tests.test_gpu_training as T
    T.test_training_steps_reduce_the_loss_and_touch_only_trainable_parameters(dev=torch.device("cpu"), steps=5)


def test_fused_adamw_on_the_emulator(emu):
    import tests.test_gpu_ops as P
    P.test_fused_adamw_matches_torch(8 * 1000 + 8, 0.0, False)
    P.test_fused_adamw_matches_torch(4099, 0.1, True)


def test_parallel_split_k_on_the_emulator(emu):
    import tests.test_gpu_gemm as G
    import tests.test_gpu_modules as M
    G.test_gemm_parallel_split_k(264, 200, 1000, 2, 0)
    G.test_gemm_group_with_parallel_split_k()
    M.test_xattn_seeded_vs_oracle(8, 128, 1, 64, 64, oracle_dt=torch.float32)       # >= 1024 rows: the dW GEMMs split K (switch on by default)
