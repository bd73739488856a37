"""A few optimisation steps of the whole model on the GPU (training.train: fwd, bwd, flat-arena AdamW)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_training_steps_reduce_the_loss_and_touch_only_trainable_parameters(dev=None, steps=12):
    import bench
    from flamingo_mini_b200.parallel import hot_path_modules
    from flamingo_mini_b200.training import train
    w = dict(bench.WORKLOADS["tiny"])
    w["lm_config"] = dict(w["lm_config"], resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)     # deterministic: same batch every step
    dev = torch.device("cuda", 0) if dev is None else dev           # the host-emulator run passes the CPU
    model = bench.build_model(w, dev)
    clip, ids, ml = bench.make_batch(w, w["B"], dev, 7, torch.bfloat16)
    frozen = {n: p.detach().clone() for n, p in model.named_parameters() if not p.requires_grad}
    before = {id(m): m._fp.ensure().clone() for m in hot_path_modules(model)}

    def batches():
        while True:
            vf = model.flamingo.resampler(clip).reshape(ids.shape[0], w["N"], 64, w["Dv"])
            yield dict(input_ids=ids, media_locations=ml, visual_features=vf, labels=ids, attention_mask=torch.ones_like(ids))

    losses = train(model, batches(), steps=steps, lr=2e-3)
    assert all(torch.isfinite(torch.tensor(losses)))
    assert min(losses[-3:]) < losses[0], losses
    for m in hot_path_modules(model):
        assert m._fp.is_attached() and not torch.equal(m._fp.flat, before[id(m)])
    for n, p in model.named_parameters():
        if not p.requires_grad:
            assert torch.equal(p, frozen[n]), n


def test_gradient_clipping_on_real_arenas(dev=None):
    """max_grad_norm over the CUDA modules' flat gradient arenas (the arena's padded tail must never contribute —
    it is allocated zeroed and no kernel writes it): the clipped total norm equals the norm over the parameters' own views."""
    import bench
    from flamingo_mini_b200.parallel import hot_path_modules
    from flamingo_mini_b200.training import train
    w = dict(bench.WORKLOADS["tiny"])
    w["lm_config"] = dict(w["lm_config"], resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    dev = torch.device("cuda", 0) if dev is None else dev
    model = bench.build_model(w, dev)
    clip, ids, ml = bench.make_batch(w, w["B"], dev, 7, torch.bfloat16)

    def batches():
        while True:
            vf = model.flamingo.resampler(clip).reshape(ids.shape[0], w["N"], 64, w["Dv"])
            yield dict(input_ids=ids, media_locations=ml, visual_features=vf, labels=ids, attention_mask=torch.ones_like(ids))

    # one plain backward: arena norm == norm over the parameter views (no stale tail), and the tail is exactly zero
    out = model(**next(batches()))
    out.loss.backward()
    for m in hot_path_modules(model):
        g = m._last_grad_arena
        views = torch.cat([v.reshape(-1) for v in m._fp.grad_views(g)])
        torch.testing.assert_close(torch.linalg.vector_norm(g), torch.linalg.vector_norm(views), rtol=1e-6, atol=0)
        used = sum(p.numel() for p in m._fp.params())
        assert used <= g.numel() < used + 8 and not g[used:].any()
    model.zero_grad(set_to_none=True)
    losses = train(model, batches(), steps=6, lr=1e-3, max_grad_norm=0.05)
    assert all(torch.isfinite(torch.tensor(losses))), losses
    assert losses[-1] < losses[0], losses
