"""LayerNorm forward/backward, text_time and cast kernels through the C ABI vs PyTorch fp32 references."""
import ctypes as C

import pytest
import torch

from flamingo_mini_b200 import _lib
from flamingo_mini_b200._lib import check
from tests._gpu_util import ptr, rel_err, stream

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("rows,D", [(37, 64), (513, 768), (100, 4096), (9, 8192)])
@pytest.mark.parametrize("x_f32", [0, 1])
def test_layernorm_fwd_bwd(rows, D, x_f32):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(rows + D)
    x = torch.randn(rows, D, device=DEV, generator=g) * 2 + 0.5
    x = x if x_f32 else x.to(torch.bfloat16)
    gamma = torch.randn(D, device=DEV, generator=g) * 0.2 + 1
    beta = torch.randn(D, device=DEV, generator=g) * 0.1
    out = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    mean = torch.empty(rows, device=DEV)
    rstd = torch.empty(rows, device=DEV)
    check(lib.fm_layernorm_fwd(ptr(x), x_f32, ptr(gamma), ptr(beta), ptr(out), 0, ptr(mean), ptr(rstd), rows, D, stream()))
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    assert rel_err(out, ref) < 5e-3                      # one bf16 rounding of the output
    torch.testing.assert_close(mean, x.float().mean(-1), rtol=1e-4, atol=1e-4)

    dy = torch.randn(rows, D, device=DEV, generator=g).to(torch.bfloat16)
    dres = torch.randn(rows, D, device=DEV, generator=g).to(torch.bfloat16)
    dx = torch.empty(rows, D, device=DEV, dtype=torch.float32)
    dgm, dbt = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
    part = torch.empty(lib.fm_layernorm_bwd_scratch_bytes(D), dtype=torch.uint8, device=DEV)
    check(lib.fm_layernorm_bwd(ptr(dy), ptr(x), x_f32, ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), 0, ptr(dx), 1,
                               ptr(dgm), ptr(dbt), ptr(part), rows, D, stream()))
    ref.backward(dy.float())
    assert rel_err(dx, xr.grad + dres.float()) < 1e-3
    assert rel_err(dgm, gr.grad) < 1e-3
    assert rel_err(dbt, br.grad) < 1e-3


def test_text_time_and_cast():
    lib = _lib.load()
    ml = (torch.rand(5, 77, device=DEV) < 0.1).to(torch.int32)
    tt = torch.empty_like(ml)
    check(lib.fm_text_time(ptr(ml), ptr(tt), 5, 77, stream()))
    assert torch.equal(tt, ml.cumsum(-1).to(torch.int32))
    src = torch.randn(100003, device=DEV)
    dst = torch.empty(100003, device=DEV, dtype=torch.bfloat16)
    check(lib.fm_cast_f32_to_bf16(ptr(src[:100000]), ptr(dst), 100000, stream()))
    assert torch.equal(dst[:100000], src[:100000].to(torch.bfloat16))


# ---- loss head (fm_cross_entropy_{fwd,bwd})
@pytest.mark.parametrize("rows,vocab,ld", [(64, 50258, 50304), (7, 1000, 1000), (33, 515, 520), (5, 8, 8), (16, 50273, 50304)])
def test_cross_entropy_vs_torch(rows, vocab, ld):
    assert _lib.has("fm_cross_entropy_fwd"), "entry point not exported by the loaded library"
    from flamingo_mini_b200 import functional as Fn
    g = torch.Generator(device=DEV).manual_seed(rows + vocab)
    logits = (torch.randn(rows, ld, device=DEV, generator=g) * 3).to(torch.bfloat16)
    logits[:, vocab:] = 77.0                                    # padding must never be read
    targets = torch.randint(0, vocab, (rows,), device=DEV, generator=g)
    targets[::5] = -100                                         # ignored rows (the last position of every sequence)
    if rows > 1:
        targets[1] = vocab - 1                                  # target in the last (partial) chunk
    x = logits.clone().requires_grad_(True)
    loss = Fn.cross_entropy(x, targets, vocab)
    ref_in = logits[:, :vocab].float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_in, targets, ignore_index=-100)
    assert abs(loss.item() - ref.item()) <= 4e-3 * abs(ref.item()) + 1e-3          # loss is returned in bf16
    (loss * 3.0).backward()
    (ref * 3.0).backward()
    assert torch.equal(x.grad[:, vocab:], torch.zeros_like(x.grad[:, vocab:]))      # padding columns: exact zeros
    assert torch.equal(x.grad[::5], torch.zeros_like(x.grad[::5]))                  # ignored rows: exact zeros
    assert rel_err(x.grad[:, :vocab], ref_in.grad) < 8e-3                           # one bf16 rounding of each entry


def test_cross_entropy_all_rows_ignored_and_graph_capture():
    assert _lib.has("fm_cross_entropy_fwd"), "entry point not exported by the loaded library"
    from flamingo_mini_b200 import functional as Fn
    logits = torch.randn(4, 64, device=DEV).to(torch.bfloat16).requires_grad_(True)
    loss = Fn.cross_entropy(logits, torch.full((4,), -100, device=DEV), 60)
    loss.backward()
    assert loss.item() == 0.0 and not logits.grad.any()
    # capturable: no host synchronisation anywhere (count and scale stay on the device)
    x = torch.randn(32, 1024, device=DEV).to(torch.bfloat16).requires_grad_(True)
    t = torch.randint(0, 1000, (32,), device=DEV)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            x.grad = None
            Fn.cross_entropy(x, t, 1000).backward()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    want = x.grad.clone()
    x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        Fn.cross_entropy(x, t, 1000).backward()
    graph.replay(); torch.cuda.synchronize()
    assert torch.equal(x.grad, want)


# ---- fused AdamW over a flat arena (fm_adamw_step) vs torch.optim.AdamW, incl. decay mask, clip scale and the bf16 shadow
@pytest.mark.parametrize("n,wd,clip", [(8 * 1000 + 8, 0.0, False), (123456, 0.1, True), (64, 0.05, False)])
def test_fused_adamw_matches_torch(n, wd, clip):
    assert _lib.has("fm_adamw_step"), "entry point not exported by the loaded library"
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(n)
    p = torch.randn(n, device=DEV, generator=g)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    shadow = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    mask = (torch.rand(n, device=DEV, generator=g) < 0.7).float()
    ref_dec = torch.nn.Parameter(p.clone()[mask.bool()])
    ref_nod = torch.nn.Parameter(p.clone()[~mask.bool()])
    opt = torch.optim.AdamW([{"params": [ref_dec], "weight_decay": wd}, {"params": [ref_nod], "weight_decay": 0.0}], lr=3e-3,
                            betas=(0.9, 0.999), eps=1e-8)
    scale = torch.tensor([0.37], device=DEV) if clip else None
    for step in range(1, 4):
        grad = torch.randn(n, device=DEV, generator=g) * (0.1 * step)
        check(lib.fm_adamw_step(ptr(p), ptr(grad), ptr(m), ptr(v), ptr(shadow), ptr(mask) if wd else None, ptr(scale), n, 3e-3, 0.9, 0.999,
                                1e-8, wd, step, stream()), "fm_adamw_step")
        gs = grad * (0.37 if clip else 1.0)
        ref_dec.grad, ref_nod.grad = gs[mask.bool()].clone(), gs[~mask.bool()].clone()
        if not wd:                       # without a mask every element is in the "decay" formula with wd = 0: same thing
            pass
        opt.step()
        want = torch.empty_like(p)
        want[mask.bool()], want[~mask.bool()] = ref_dec.detach(), ref_nod.detach()
        torch.testing.assert_close(p, want, rtol=2e-5, atol=2e-6)
        assert torch.equal(shadow, p.to(torch.bfloat16))
