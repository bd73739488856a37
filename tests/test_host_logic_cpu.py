"""CPU-only tests of the host-side logic around the kernels: flat parameter storage, cached-kv views, layouts,
config defaults, processor text logic, generation plumbing and the aligned loss head."""
import os

import pytest
import torch
import torch.nn.functional as F

from flamingo_mini_b200 import GatedCrossAttentionBlock, PerceiverResampler
from flamingo_mini_b200 import functional as Fn
from flamingo_mini_b200.configuration_flamingo import FlamingoConfig
from flamingo_mini_b200.flamingo_processor import FlamingoProcessor
from flamingo_mini_b200.gated_cross_attention import _kv_buffer, _kv_views
from flamingo_mini_b200.modeling_flamingo import FlamingoBaseModel, FlamingoModel
from flamingo_mini_b200.utils import get_common_prefix_length


def test_flat_params_views_and_reattach():
    blk = GatedCrossAttentionBlock(dim=128, dim_visual=64)
    fp = blk._fp
    assert not fp.is_attached()
    before = {n: p.detach().clone() for n, p in blk.named_parameters()}
    flat = fp.ensure()
    assert fp.is_attached() and flat.numel() == fp.total and fp.total % 8 == 0
    for n, p in blk.named_parameters():                      # values survive the move into the flat buffer
        assert torch.equal(p, before[n])
    with torch.no_grad():                                    # an in-place update of a parameter is visible in the flat buffer
        blk.attn.to_q.weight.add_(1.0)
    off = dict((id(p), o) for p, o in fp.slots)[id(blk.attn.to_q.weight)]
    assert torch.equal(flat[off:off + 512 * 128].view(512, 128), blk.attn.to_q.weight)
    blk.double().float()                                     # nn.Module._apply re-creates parameter storage
    assert not fp.is_attached()
    fp.ensure()
    assert fp.is_attached()
    g = torch.arange(fp.total, dtype=torch.float32)
    views = fp.grad_views(g)
    assert [v.shape for v in views] == [p.shape for p in fp.params()]
    assert views[2].data_ptr() == g.data_ptr() + 4 * off     # slot order == registration order (to_q is third)


def test_layout_offsets_are_aligned_and_disjoint():
    L = Fn.xattn_layout(768, 1024, 8, 64, 3072)
    sizes = dict(attn_norm_w=768, attn_norm_b=768, to_q=512 * 768, to_kv=1024 * 1024, to_out=768 * 512, ffw_norm_w=768,
                 ffw_norm_b=768, ffw_w1=3072 * 768, ffw_w2=768 * 3072, alpha_attn=1, alpha_ffw=1)
    spans = sorted((getattr(L, k), getattr(L, k) + n) for k, n in sizes.items())
    assert all(a2 >= b1 for (_, b1), (a2, _) in zip(spans, spans[1:])) and spans[-1][1] <= L.total
    for k in ("to_q", "to_kv", "to_out", "ffw_w1", "ffw_w2"):
        assert getattr(L, k) % 8 == 0                          # bf16 shadow operands must be 16-byte aligned for TMA
    R = Fn.resampler_layout(1024, 6, 8, 64, 64, 4, 4096)
    assert R.layer0 % 8 == 0 and R.layer_stride % 8 == 0 and R.to_q % 8 == 0 and R.to_out % 8 == 0 and R.ffw_w1 % 8 == 0
    res = PerceiverResampler(dim=64, depth=2)
    offs = sorted((o, o + p.numel()) for p, o in res._fp.slots)
    assert all(a2 >= b1 for (_, b1), (a2, _) in zip(offs, offs[1:])) and offs[-1][1] <= res._fp.total


def test_kv_views_roundtrip():
    B, H, V, dh = 2, 8, 128, 64
    kv = torch.randn(B * V, 2 * H * dh).to(torch.bfloat16)
    k, v = _kv_views(kv, B, H, dh)
    assert k.shape == (B, H, V, dh) and v.shape == (B, H, V, dh)
    assert torch.equal(k[1, 3, 5], kv[1 * V + 5, 3 * dh:4 * dh]) and torch.equal(v[0, 7, 9], kv[9, 512 + 7 * dh:512 + 8 * dh])
    back = _kv_buffer(k, v)
    assert back.data_ptr() == kv.data_ptr() and torch.equal(back, kv)          # our own views: no copy
    back2 = _kv_buffer(k.float().contiguous(), v.float().contiguous())          # foreign tensors (e.g. after beam reorder)
    assert back2.dtype == torch.bfloat16 and torch.equal(back2, kv)
    sel = torch.tensor([1, 0])
    back3 = _kv_buffer(k.index_select(0, sel), v.index_select(0, sel))
    assert torch.equal(back3.view(B, V, -1), kv.view(B, V, -1)[sel])


def test_config_defaults_match_reference():
    c = FlamingoConfig()
    want = dict(lm="gpt2", clip_model_type="openai/clip-vit-base-patch32", dim=1024, dim_visual=768, xattn_every=1,
                xattn_dim_head=64, xattn_heads=8, xattn_ff_mult=4, xattn_act="gelu", resampler_depth=6,
                resampler_dim_head=64, resampler_heads=8, resampler_num_latents=64, resampler_num_time_embeds=4,
                resampler_ff_mult=4, resampler_act="gelu", freeze_language_model=True, freeze_vision_model=True)
    for k, v in want.items():
        assert getattr(c, k) == v, k
    d = FlamingoConfig(dim=768, xattn_act="sqrelu").to_dict()
    assert d["dim"] == 768 and d["xattn_act"] == "sqrelu"
    ref = "/root/reference/flamingo_mini/configuration_flamingo.py"
    if os.path.exists(ref):                                   # build container only: every reference field is present
        import re
        for name in re.findall(r"self\.(\w+) = \1", open(ref).read()):
            assert hasattr(c, name), name


class _FakeTok:
    eos_token = "</s>"
    pad_token = None
    add_bos_token = False
    vocab = {"<": 27, " <": 1279}

    def add_tokens(self, t):
        self.added = t

    def encode(self, s):
        return [self.vocab.get(s, 5)]

    def __call__(self, text, **kw):
        class R:
            pass
        r = R()
        r.input_ids = torch.tensor([[50, 27, 9, 9, 1279, 4], [27, 3, 3, 3, 3, 3]])
        r.attention_mask = torch.ones_like(r.input_ids)
        return r


def test_processor_text_logic_with_fake_tokenizer():
    proc = FlamingoProcessor(FlamingoConfig(), tokenizer=_FakeTok(), vision_processor=object())
    assert proc.leq_ids == [27, 1279] and proc.tokenizer.pad_token == "</s>" and proc.tokenizer.added == "<EOC>"
    ids, ml, mask = proc.encode_text(["a", "b"])
    assert ml.dtype == torch.int64 and ml.tolist() == [[0, 1, 0, 0, 1, 0], [1, 0, 0, 0, 0, 0]]
    assert proc.prepare_caption("a cat") == "<image>a cat<EOC></s>"
    assert proc.remove_tags(["<image>a cat<EOC></s>", " x "]) == ["a cat", "x"]


def test_generation_plumbing_and_prefix_length():
    prep = FlamingoModel.prepare_inputs_for_generation
    ids = torch.arange(12).view(4, 3)
    ml = torch.tensor([[1, 0, 0], [0, 1, 0]])
    vf = torch.randn(2, 1, 64, 8)
    out = prep(None, ids, media_locations=ml, visual_features=vf, attention_mask=None)
    assert out["media_locations"].tolist() == [[1, 0, 0], [1, 0, 0], [0, 1, 0], [0, 1, 0]]      # beams expand consecutively
    assert torch.equal(out["visual_features"][1], vf[0]) and out["input_ids"].shape == (4, 3)
    out = prep(None, ids, media_locations=ml, past_key_values=("x", "y"))
    assert out["input_ids"].shape == (4, 1) and out["past_key_values"] == ("x", "y")
    xattn_past = ((torch.arange(4.0).view(4, 1, 1, 1), torch.arange(4.0).view(4, 1, 1, 1)),)
    lm_past = ((torch.arange(4.0).view(4, 1), torch.arange(4.0).view(4, 1)),)
    xb, lb = FlamingoModel._reorder_cache(None, (xattn_past, lm_past), torch.tensor([3, 3, 0, 1]))
    assert xb[0][0].flatten().tolist() == [3, 3, 0, 1] and lb[0][1].flatten().tolist() == [3, 3, 0, 1]
    assert get_common_prefix_length(torch.tensor([[1, 2, 3, 4], [1, 2, 9, 4], [1, 2, 3, 5]])) == 2
    assert get_common_prefix_length(torch.tensor([[1, 2], [1, 2]])) == 2


def test_aligned_loss_head_math_is_identical():
    """Padded logits with -inf bias + ignore_index == plain shifted cross-entropy (value and gradient)."""
    torch.manual_seed(0)
    V, D = 101, 16
    w = torch.randn(V, D, requires_grad=True)
    h = torch.randn(2, 7, D)
    labels = torch.randint(0, V, (2, 7))
    pad = (-V) % 64
    bias = torch.zeros(V + pad)
    bias[V:] = float("-inf")
    padded = F.linear(h, F.pad(w, (0, 0, 0, pad)), bias)
    loss_fast = FlamingoBaseModel._shifted_cross_entropy(padded[..., :V], padded, labels, "mean")
    loss_ref = FlamingoBaseModel._shifted_cross_entropy(F.linear(h, w), None, labels, "mean")
    assert torch.allclose(loss_fast, loss_ref, atol=1e-6)
    g1, = torch.autograd.grad(loss_fast, w, retain_graph=True)
    g0, = torch.autograd.grad(loss_ref, w)
    assert torch.allclose(g0, g1, atol=1e-6)
    per_tok = FlamingoBaseModel._shifted_cross_entropy(padded[..., :V], padded, labels, "none")    # falls back to the plain path
    assert per_tok.shape == (2 * 6,)


def test_unsupported_configurations_are_rejected():
    with pytest.raises(Exception):
        GatedCrossAttentionBlock(dim=128, dim_visual=64, dim_head=32)        # attention cores are specialised for 64-wide heads
    with pytest.raises(ValueError):
        GatedCrossAttentionBlock(dim=128, dim_visual=64, n_visual=32)
    with pytest.raises(Exception):
        PerceiverResampler(dim=100, depth=1)                                   # width must be a multiple of 64
    with pytest.raises(AssertionError):
        GatedCrossAttentionBlock(dim=128, dim_visual=64, act="swish")


def test_fused_gelu_swap_is_the_same_function():
    """lm_fused_gelu: the frozen GPT-2's gelu_new modules become nn.GELU('tanh'); logits and LM-input gradients agree with
    the un-fused model to fp32 round-off, no parameter / checkpoint key changes."""
    from flamingo_mini_b200.configuration_flamingo import FlamingoConfig
    from flamingo_mini_b200.modeling_flamingo import FlamingoGPT2
    from transformers.activations import NewGELUActivation
    lm_cfg = dict(n_embd=64, n_layer=2, n_head=2, vocab_size=53, n_positions=32, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    clip_cfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, image_size=32, patch_size=16)
    models = []
    for fused in (True, False):
        torch.manual_seed(0)
        models.append(FlamingoGPT2(FlamingoConfig(lm="gpt2", dim=64, dim_visual=64, lm_config=lm_cfg, clip_config=clip_cfg,
                                                  lm_fused_gelu=fused)))
    a, b = models
    assert not any(isinstance(m, NewGELUActivation) for m in a.lm.modules())
    assert sum(isinstance(m, NewGELUActivation) for m in b.lm.modules()) == 2
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 9, 64)
    for la, lb in zip(a.lm.h, b.lm.h):          # the spliced blocks need CUDA; the swap only touches the frozen MLPs
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya, yb = la.lm_block.mlp(xa), lb.lm_block.mlp(xb)
        torch.testing.assert_close(ya, yb, rtol=1e-5, atol=1e-6)
        ga, gb = torch.autograd.grad(ya.square().sum(), xa)[0], torch.autograd.grad(yb.square().sum(), xb)[0]
        torch.testing.assert_close(ga, gb, rtol=1e-4, atol=1e-6)
