"""A few optimisation steps of the whole model on the GPU (training.train: fwd, bwd, flat-arena AdamW)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.first_hw_run
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
