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
