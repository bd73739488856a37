"""Host-side training driver (flamingo_mini_b200/training.py) on CPU: the flat-arena AdamW against torch.optim.AdamW,
trainable-only checkpoints, and the loop itself with the oracle modules standing in for the CUDA ones."""
import os

import torch

from flamingo_mini_b200 import GatedCrossAttentionBlock, PerceiverResampler
from flamingo_mini_b200.configuration_flamingo import FlamingoConfig
from flamingo_mini_b200.modeling_flamingo import FlamingoModel
from flamingo_mini_b200.training import ArenaAdamW, DataCollator, constant_schedule_with_warmup, load_trainable, save_trainable, train
from oracle.oracle_modules import swap_in_oracle


def _fake_backward(mod, seed):
    g = torch.Generator().manual_seed(seed)
    arena = torch.randn(mod._fp.total, generator=g)
    mod._last_grad_arena = arena
    for p, v in zip(mod._fp.params(), mod._fp.grad_views(arena)):
        p.grad = v
    return arena


def test_arena_adamw_matches_torch_adamw():
    for wd in (0.0, 0.1):
        torch.manual_seed(0)
        mods = [GatedCrossAttentionBlock(dim=64, dim_visual=64), PerceiverResampler(dim=64, depth=1)]
        for m in mods:
            m._fp.attach()
            with torch.no_grad():
                m._fp.flat.copy_(torch.randn(m._fp.total) * 0.3)
        extra = torch.nn.Parameter(torch.randn(7, 5))
        # per-tensor reference with the decay grouping of the INSTALLED HF Trainer (get_decay_parameter_names): everything except
        # nn.LayerNorm parameters and "bias"/"norm" names is decayed - gates, latents and time_pos_emb included
        from transformers import Trainer
        ref_params = [[(n, torch.nn.Parameter(p.detach().clone())) for n, p in m.named_parameters()] for m in mods]
        ref_extra = torch.nn.Parameter(extra.detach().clone())
        hf_decay = [set(Trainer.get_decay_parameter_names(None, m)) for m in mods]
        assert "alpha_attn" in hf_decay[0] and "latents" in hf_decay[1] and "attn.norm.weight" not in hf_decay[0]
        groups = [{"params": [p for i, ps in enumerate(ref_params) for n, p in ps if n in hf_decay[i]] + [ref_extra], "weight_decay": wd},
                  {"params": [p for i, ps in enumerate(ref_params) for n, p in ps if n not in hf_decay[i]], "weight_decay": 0.0}]
        ref_opt = torch.optim.AdamW(groups, lr=1e-2)
        opt = ArenaAdamW(mods, [extra], lr=1e-2, weight_decay=wd)
        for step in range(3):
            for i, m in enumerate(mods):
                _fake_backward(m, 10 * step + i)
                names = [n for n, _ in m.named_parameters()]
                for (n, rp), p in zip(ref_params[i], [dict(m.named_parameters())[n] for n in names]):
                    rp.grad = p.grad.detach().clone()
            extra.grad = torch.full_like(extra, 0.1 * (step + 1))
            ref_extra.grad = extra.grad.clone()
            opt.step()
            ref_opt.step()
            for i, m in enumerate(mods):
                assert m._fp.is_attached() and m._fp._shadow_ver is None
                for (n, rp) in ref_params[i]:
                    torch.testing.assert_close(dict(m.named_parameters())[n].detach(), rp.detach(), rtol=1e-5, atol=1e-6, msg=lambda s: f"{n}: {s}")
            torch.testing.assert_close(extra.detach(), ref_extra.detach(), rtol=1e-5, atol=1e-6)
        opt.zero_grad()
        assert all(p.grad is None for m in mods for p in m.parameters()) and mods[0]._last_grad_arena is None


def test_arena_adamw_uses_the_accumulated_gradient():
    """Two backwards without zero_grad in between: the second writes a temporary arena that autograd ADDS into .grad (views of
    the module's cached arena).  The optimizer must step on the accumulated sum, not on the arena the last backward wrote."""
    torch.manual_seed(0)
    m = GatedCrossAttentionBlock(dim=64, dim_visual=64)
    m._fp.attach()
    ref = [torch.nn.Parameter(p.detach().clone()) for p in m.parameters()]
    ref_opt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.0)
    opt = ArenaAdamW([m], lr=1e-2)
    for step in range(2):
        opt.zero_grad()
        for micro in range(2):                                        # what _XattnFn.backward + AccumulateGrad do
            g = m._fp.grad_arena(m)
            g.copy_(torch.randn(m._fp.total, generator=torch.Generator().manual_seed(10 * step + micro)))
            m._last_grad_arena = g
            for p, v in zip(m._fp.params(), m._fp.grad_views(g)):
                if p.grad is None:
                    p.grad = v
                else:
                    p.grad.add_(v)
        assert m._last_grad_arena is not m._grad_arena                # the last backward's arena holds only the second micro-batch
        flat, aliased = m._fp.current_grad(m)
        assert aliased and flat is m._grad_arena
        for rp, p in zip(ref, m.parameters()):
            rp.grad = p.grad.detach().clone()
        opt.step(); ref_opt.step()
        for rp, (n, p) in zip(ref, m.named_parameters()):
            torch.testing.assert_close(p.detach(), rp.detach(), rtol=1e-5, atol=1e-6, msg=lambda s: f"{n}: {s}")


def test_schedule_and_collator():
    lr = constant_schedule_with_warmup(1e-3, 4)
    assert [round(lr(s) / 1e-3, 2) for s in range(6)] == [0.0, 0.25, 0.5, 0.75, 1.0, 1.0]
    from transformers import get_constant_schedule_with_warmup           # the scheduler training/train.py:165 names
    sgd = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
    hf = get_constant_schedule_with_warmup(sgd, 4)
    for s in range(6):
        assert abs(sgd.param_groups[0]["lr"] - lr(s)) < 1e-12
        sgd.step(); hf.step()

    class _Proc:
        def __call__(self, text):
            n = len(text)
            return dict(input_ids=torch.arange(n * 3).reshape(n, 3), media_locations=torch.zeros(n, 3, dtype=torch.long),
                        attention_mask=torch.ones(n, 3, dtype=torch.long))

    out = DataCollator(_Proc())([(torch.zeros(3, 4, 4), "<image>a cat"), (torch.ones(3, 4, 4), "<image>a dog")])
    assert out["pixel_values"].shape == (2, 3, 4, 4) and torch.equal(out["labels"], out["input_ids"])
    assert set(out) == {"pixel_values", "labels", "input_ids", "media_locations", "attention_mask"}


def test_train_loop_and_trainable_checkpoint(golden_dir, tmp_path):
    fx = torch.load(os.path.join(golden_dir, "model_opt_tiny.pt"))
    cfg = FlamingoConfig(lm="facebook/opt-125m", dim=64, dim_visual=64, xattn_every=1, resampler_depth=1,
                         lm_config=fx["opt_cfg"], clip_config=fx["clip_cfg"])
    model = FlamingoModel(cfg)
    model.load_state_dict(fx["state_dict"], strict=True)
    swap_in_oracle(model, copy_weights=True)        # CPU stand-ins for the CUDA modules: the loop itself is under test
    with torch.no_grad():
        for layer in model.flamingo.get_modified_layers():
            layer.xattn_block.alpha_attn.fill_(0.3); layer.xattn_block.alpha_ffw.fill_(0.3)
    ids = fx["input_ids"]
    batch = dict(input_ids=ids, media_locations=fx["media_locations"], pixel_values=fx["pixel_values"], labels=ids,
                 attention_mask=torch.ones_like(ids))
    frozen_before = {n: p.detach().clone() for n, p in model.named_parameters() if not p.requires_grad}
    torch.manual_seed(0)
    losses = train(model, iter([batch] * 8), steps=8, lr=3e-3, warmup_steps=2, max_grad_norm=1.0)
    assert losses[-1] < losses[0]                   # the same batch 8 times: the loss must go down
    for n, p in model.named_parameters():
        if not p.requires_grad:
            assert torch.equal(p, frozen_before[n]), f"frozen parameter {n} changed"
    path = str(tmp_path / "ck.pt")
    save_trainable(model, path, step=8)
    ck = torch.load(path)
    assert sorted(ck["trainable"]) == sorted(model.flamingo.state_dict_trainable()) and ck["step"] == 8
    fresh = FlamingoModel(cfg)
    fresh.load_state_dict(fx["state_dict"], strict=True)
    swap_in_oracle(fresh, copy_weights=True)
    assert load_trainable(fresh, path) == 8
    a, b = model.flamingo.state_dict_trainable(), fresh.flamingo.state_dict_trainable()
    for k in a:
        torch.testing.assert_close(a[k], b[k])
