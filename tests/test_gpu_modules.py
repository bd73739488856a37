"""Parity of the CUDA modules (through the reference's nn.Module API and the C ABI) against the CPU oracle.

Tolerances (stated per north_star: bf16 tensor-core operands, fp32 accumulation / statistics):
  forward outputs   relative L2 error <= 1.5e-2 vs the fp64 oracle
  gradients         relative L2 error <= 4e-2   vs the fp64 oracle (per tensor; tensors whose oracle norm is ~0 are
                    compared on absolute error)
Golden fixtures (tests/golden/*.pt) were produced by the unmodified reference; seeded larger cases use the oracle
computed on the host in the same test.
"""
import os

import pytest
import torch

from flamingo_mini_b200 import GatedCrossAttentionBlock, PerceiverResampler
from oracle import flamingo_oracle as O
from tests._gpu_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
FWD_TOL, BWD_TOL = 1.5e-2, 6e-2   # 6e-2: tiny (width-64) golden cases sit on ReLU/sqReLU kinks; bf16 P/dS operands


def _close(got, want, tol, name, elementwise=True):
    """whole-tensor relative L2 error <= tol AND element-wise |got - want| <= 4*tol*rms(want) + tol*|want| (a single wrong
    row / tile / element cannot hide inside a good norm).  elementwise=False only for gradients behind a ReLU: its
    derivative is a step, so ONE pre-activation within bf16 rounding of zero legitimately flips a whole row of dW."""
    want = want.to(got.device)
    n = want.float().norm().item()
    if n < 1e-6:
        assert got.float().norm().item() < 1e-3, f"{name}: expected ~0, got norm {got.float().norm().item()}"
        return
    e = rel_err(got, want)
    assert e <= tol, f"{name}: rel L2 err {e:.3e} > {tol}"
    if want.numel() > 1 and elementwise:
        rms = n / want.numel() ** 0.5
        torch.testing.assert_close(got.float(), want.float().reshape(got.shape), rtol=tol, atol=4 * tol * rms,
                                   msg=lambda m: f"{name}: {m}")


def _golden_param_grads(fx, oracle_fn, params, inputs, got, tol, elementwise=True):
    """Parameter gradients of a golden case.  Small tensors are stored in full by the reference run; the large ones (every weight
    matrix) are stored as an order-sensitive digest (sum, norm, two seeded random projections).  For those the oracle's FULL
    gradient is recomputed here, checked against the reference's digest to 1e-6 (so it IS the reference's tensor up to fp64
    round-off, transpositions and permutations included), and the CUDA gradient is then compared with it element by element."""
    from tests.golden.make_golden import digest
    _, _, o_gp = _oracle_grads(oracle_fn, inputs, params, fx["cot"])
    for n, ref in fx["dparams"].items():
        if "full" in ref:
            _close(got[n], ref["full"], tol, n, elementwise)
        else:
            torch.testing.assert_close(digest(n, o_gp[n]), ref["digest"], rtol=1e-6, atol=1e-6, msg=lambda m: f"oracle vs reference digest {n}: {m}")
            _close(got[n], o_gp[n].reshape(got[n].shape), tol, n, elementwise)
            d = digest(n, got[n].detach().double().cpu())
            for j in (2, 3):       # the stored projections themselves: error of a projection ~ N(0, |err|^2)
                assert abs(d[j] - ref["digest"][j]).item() <= 4 * tol * ref["digest"][1].item(), f"{n}: projection {j - 2} off"


def _oracle_grads(fn, inputs, params, cot, dt=torch.float64):
    """run oracle in fp64 (fp32 for the full-size cases) on CPU; returns out, input grads, param grads"""
    p = {k: v.detach().to(dt).cpu().requires_grad_(True) for k, v in params.items()}
    ins = [None if t is None else (t.detach().to(dt).cpu().requires_grad_(True) if t.is_floating_point() else t.cpu())
           for t in inputs]
    out = fn(ins, p)
    out.backward(cot.to(dt).cpu())
    return out.detach(), [None if (t is None or not t.is_floating_point()) else t.grad for t in ins], {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("name", ["res_img", "res_vid", "res_relu", "res_h4"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resampler_golden(golden_dir, name, dtype):
    fx = torch.load(os.path.join(golden_dir, f"{name}.pt"))
    c = fx["case"]
    heads, ff_mult = c.get("heads", 8), c.get("ff_mult", 4)
    params = O.seeded_params(O.resampler_param_shapes(c["dim"], c["depth"], heads=heads, ff_mult=ff_mult), c["seed"])
    m = PerceiverResampler(dim=c["dim"], depth=c["depth"], heads=heads, ff_mult=ff_mult, act=c["act"])
    m.load_state_dict(params, strict=True)
    m = m.to(DEV)
    x = fx["x"].to(DEV).to(dtype)
    out = m(x)
    assert out.dtype == dtype
    tol = FWD_TOL if dtype == torch.float32 else 2.5e-2          # bf16 inputs are themselves rounded
    _close(out, fx["out"], tol, "out")
    out.backward(fx["cot"].to(DEV).to(dtype))
    full = {n: p.grad for n, p in m.named_parameters()}
    params64 = O.seeded_params(O.resampler_param_shapes(c["dim"], c["depth"], heads=heads, ff_mult=ff_mult), c["seed"], dtype=torch.float64)
    _golden_param_grads(fx, lambda i, p: O.perceiver_resampler(i[0], p, c["depth"], heads=heads, act=c["act"]), params64, [fx["x"]], full,
                        BWD_TOL if dtype == torch.float32 else 6e-2, elementwise=c["act"] != "relu")


@pytest.mark.parametrize("name", ["xattn_edge", "xattn_sq", "xattn_h2"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_xattn_golden(golden_dir, name, dtype):
    fx = torch.load(os.path.join(golden_dir, f"{name}.pt"))
    c = fx["case"]
    heads, ff_mult = c.get("heads", 8), c.get("ff_mult", 4)
    params = O.seeded_params(O.xattn_param_shapes(c["dim"], c["dim_visual"], heads=heads, ff_mult=ff_mult), c["seed"])
    m = GatedCrossAttentionBlock(dim=c["dim"], dim_visual=c["dim_visual"], heads=heads, ff_mult=ff_mult, act=c["act"])
    m.load_state_dict(params, strict=True)
    m = m.to(DEV)
    y = fx["y"].to(DEV).to(dtype).requires_grad_(True)
    vis = fx["vis"].to(DEV).requires_grad_(True)
    ml = fx["media_locations"].to(DEV)
    out, (k, v) = m(y, vis, ml, output_kv=True)
    tol = FWD_TOL if dtype == torch.float32 else 2.5e-2
    _close(out, fx["out"], tol, "out")
    _close(k, fx["k"], 1e-2, "k")
    _close(v, fx["v"], 1e-2, "v")
    out.backward(fx["cot"].to(DEV).to(dtype))
    btol = BWD_TOL if dtype == torch.float32 else 6e-2
    _close(y.grad, fx["dy"], btol, "dy")
    _close(vis.grad, fx["dvis"], btol, "dvis")
    full = {n: p.grad for n, p in m.named_parameters()}
    params64 = O.seeded_params(O.xattn_param_shapes(c["dim"], c["dim_visual"], heads=heads, ff_mult=ff_mult), c["seed"], dtype=torch.float64)
    _golden_param_grads(fx, lambda i, p: O.gated_xattn_block(i[0], i[1], i[2], p, heads=heads, act=c["act"])[0], params64,
                        [fx["y"], fx["vis"], fx["media_locations"]], full, btol)
    # cached decode path: last 3 tokens with previous_kv
    with torch.no_grad():
        oc, _ = m(y[:, -3:].detach(), None, ml, previous_kv=(k.detach(), v.detach()))
    _close(oc, fx["out_cached_last3"], tol, "cached")


def test_xattn_identity_at_zero_gate():
    """reference: alpha = 0 (its init) => block output torch.equal to the input."""
    m = GatedCrossAttentionBlock(dim=256, dim_visual=128).to(DEV)
    for dtype in (torch.bfloat16, torch.float32):
        y = torch.randn(2, 40, 256, device=DEV).to(dtype)
        vis = torch.randn(2, 1, 64, 128, device=DEV)
        ml = torch.zeros(2, 40, dtype=torch.long, device=DEV); ml[:, 0] = 1
        out, kv = m(y, vis, ml)
        assert kv is None and torch.equal(out, y)


def test_parameters_cast_before_the_first_forward():
    """`model.lm.to(torch.bfloat16)` also reaches the xattn blocks inside the LM layers.  The modules keep fp32 master weights:
    they re-attach to their flat fp32 buffer before autograd sees the parameters, so already the FIRST backward hands out fp32
    gradients that are views of the module's arena (not per-parameter bf16 copies made by the engine)."""
    m = GatedCrossAttentionBlock(dim=128, dim_visual=64).to(DEV).to(torch.bfloat16)
    r = PerceiverResampler(dim=64, depth=1).to(DEV).to(torch.bfloat16)
    with torch.no_grad():
        m.alpha_attn.fill_(0.5); m.alpha_ffw.fill_(0.5)
    x = torch.randn(2, 1, 9, 64, device=DEV).to(torch.bfloat16)
    vis = r(x).reshape(2, 1, 64, 64)
    ml = torch.zeros(2, 24, dtype=torch.long, device=DEV); ml[:, 0] = 1
    out, _ = m(torch.randn(2, 24, 128, device=DEV).to(torch.bfloat16), vis, ml)
    out.float().square().mean().backward()
    for mod in (m, r):
        arena = mod._last_grad_arena
        lo, hi = arena.data_ptr(), arena.data_ptr() + 4 * arena.numel()
        for n, p in mod.named_parameters():
            assert p.dtype == torch.float32 and p.grad is not None and p.grad.dtype == torch.float32, n
            assert lo <= p.grad.data_ptr() < hi, f"{n}: gradient is a copy, not a view of the arena"
            assert torch.isfinite(p.grad).all(), n


@pytest.mark.parametrize("B,S,N,D,Dv", [(3, 200, 2, 256, 192), (2, 128, 1, 768, 768),
                                        (2, 256, 1, 1280, 1024),      # C3-shaped: gpt2-large width (5 x 256-column tiles)
                                        (1, 256, 4, 2048, 1024),      # C4-shaped: opt-1.3b width, 4 images
                                        (1, 128, 1, 4096, 1024)])     # C5-shaped: opt-6.7b width
def test_xattn_seeded_vs_oracle(B, S, N, D, Dv, heads=8, gate_tol=5e-2, oracle_dt=torch.float64):
    params = O.seeded_params(O.xattn_param_shapes(D, Dv, heads=heads), 123)
    m = GatedCrossAttentionBlock(dim=D, dim_visual=Dv, heads=heads)
    m.load_state_dict(params); m = m.to(DEV)
    g = torch.Generator().manual_seed(9)
    y = torch.randn(B, S, D, generator=g).to(torch.bfloat16)
    vis = torch.randn(B, N, 64, Dv, generator=g).to(torch.bfloat16)
    ml = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        for j in range(N):
            ml[b, (j * S) // N + (b % 3)] = 1
    cot = torch.randn(B, S, D, generator=g).to(torch.bfloat16)
    o_out, o_gin, o_gp = _oracle_grads(lambda i, p: O.gated_xattn_block(i[0], i[1], i[2], p, heads=heads)[0], [y, vis, ml], params, cot,
                                       dt=oracle_dt)
    yd, vd = y.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    out, _ = m(yd, vd, ml.to(DEV))
    _close(out, o_out, 2e-2, "out")
    out.backward(cot.to(DEV))
    _close(yd.grad, o_gin[0], 5e-2, "dy")
    _close(vd.grad, o_gin[1], 5e-2, "dvis")
    for n, p in m.named_parameters():
        # the two gate gradients are scalars: sums of B*S*D signed bf16 products, so their error is a random walk in the
        # summands, not a fraction of the (cancelling) total; callers with few summands per unit of total pass gate_tol
        _close(p.grad, o_gp[n], gate_tol if n.startswith("alpha_") else 5e-2, n)


@pytest.mark.parametrize("BN,T,F,Dv,depth", [(4, 1, 50, 256, 2), (2, 2, 33, 128, 1), (3, 1, 257, 128, 1),
                                             (2, 1, 257, 1024, 1)])    # ViT-L/14 width
def test_resampler_seeded_vs_oracle(BN, T, F, Dv, depth, heads=8, oracle_dt=torch.float64):
    params = O.seeded_params(O.resampler_param_shapes(Dv, depth, heads=heads), 321)
    m = PerceiverResampler(dim=Dv, depth=depth, heads=heads)
    m.load_state_dict(params); m = m.to(DEV)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(BN, T, F, Dv, generator=g).to(torch.bfloat16)
    cot = torch.randn(BN, 64, Dv, generator=g).to(torch.bfloat16)
    o_out, _, o_gp = _oracle_grads(lambda i, p: O.perceiver_resampler(i[0], p, depth, heads=heads), [x], params, cot, dt=oracle_dt)
    out = m(x.to(DEV))
    _close(out, o_out, 2e-2, "out")
    out.backward(cot.to(DEV))
    for n, p in m.named_parameters():
        _close(p.grad, o_gp[n], 6e-2, n)


def test_full_size_c2_block_and_resampler():
    """BASELINE.json configs[1] at its real size: one gated xattn block at B=32, S=128, D=Dv=768 (M = 4096 rows: multi-wave
    persistent GEMM grids, every tile-width choice of the benchmark) and the resampler on 32 images x 50 CLIP tokens (depth 2 of
    the benchmark's 6 keeps the fp32 CPU oracle to a few seconds)."""
    test_xattn_seeded_vs_oracle(32, 128, 1, 768, 768, oracle_dt=torch.float32)
    test_resampler_seeded_vs_oracle(32, 1, 50, 768, 2, oracle_dt=torch.float32)


def test_weight_gradients_with_parallel_split_k():
    """>= 1024 rows at narrow widths: the weight-gradient GEMMs (few output tiles, long K) take the parallel split-K path
    (FM_OPT_DW_SPLITK: zeroed gradient + TMA reduce-add per K range), single and grouped; with the switch off the plain path."""
    from tests._gpu_util import set_option
    for on in (1, 0):
        assert set_option("dw_splitk", on)
        try:
            test_xattn_seeded_vs_oracle(8, 128, 1, 64, 64, oracle_dt=torch.float32)
            test_resampler_seeded_vs_oracle(16, 1, 10, 64, 1, oracle_dt=torch.float32)
        finally:
            set_option("dw_splitk", 1)


def _head_counts_supported() -> bool:
    """The library takes 1..64 heads of width 64 (round 1's first library was specialised for 8)."""
    from flamingo_mini_b200 import functional as Fn
    try:
        Fn.xattn_layout(128, 64, 4, 64, 512)
        return True
    except RuntimeError:
        return False


@pytest.mark.parametrize("heads", [1, 4, 12])
def test_modules_with_other_head_counts(heads):
    """heads is a constructor argument of the reference modules (gated_cross_attention.py:16-24, perceiver_resampler.py:100-111)."""
    assert _head_counts_supported(), "the library must take 1..64 heads of width 64"
    test_xattn_seeded_vs_oracle(2, 130, 2, 128, 64, heads=heads, gate_tol=0.15)
    test_resampler_seeded_vs_oracle(2, 1, 50, 128, 1, heads=heads)
    blk = GatedCrossAttentionBlock(dim=128, dim_visual=64, heads=heads).to(DEV)          # cached decoding keeps the (B,H,V,dh) contract
    with torch.no_grad():
        y = torch.randn(2, 9, 128, device=DEV).to(torch.bfloat16)
        vis = torch.randn(2, 1, 64, 64, device=DEV).to(torch.bfloat16)
        ml = torch.zeros(2, 9, dtype=torch.long, device=DEV); ml[:, 0] = 1
        blk.alpha_attn.fill_(0.7); blk.alpha_ffw.fill_(0.3)
        full, (k, v) = blk(y, vis, ml, output_kv=True)
        assert k.shape == (2, heads, 64, 64) and v.shape == k.shape
        last, _ = blk(y[:, -1:], None, ml, previous_kv=(k, v))
        _close(last, full[:, -1:].float().cpu(), 2e-2, "cached step")


def test_resampler_rejects_too_many_frames():
    m = PerceiverResampler(dim=64, depth=1).to(DEV)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 5, 3, 64, device=DEV))


# ---- scheduling switches (fm_set_option): none of them may change a result beyond summation order.  Every switch is
# swept one at a time, plus programmatic dependent launch on / off with the side stream off.
OPTION_SETS = [
    dict(side_stream=0), dict(gemm_group=0), dict(epi_prefetch=1), dict(alpha_from_dw2=0), dict(ln_reduce_side=0), dict(pdl=1),
    dict(dattn_from_gemm=0), dict(attn_tmem_compact=0), dict(pdl=0), dict(dw_splitk=0),
    dict(pdl=1, side_stream=0),
    dict(side_stream=0, gemm_group=0, epi_prefetch=0, alpha_from_dw2=0, ln_reduce_side=0, pdl=0, dattn_from_gemm=0, attn_tmem_compact=0),
]
OPTION_DEFAULTS = dict(side_stream=1, gemm_group=1, epi_prefetch=0, alpha_from_dw2=1, pdl=1, ln_reduce_side=1, dattn_from_gemm=1, attn_tmem_compact=1,
                       defer_join=0, dw_splitk=1)


@pytest.mark.parametrize("opts", OPTION_SETS, ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
def test_modules_under_scheduling_options(opts):
    from tests._gpu_util import set_option
    assert all(set_option(k, OPTION_DEFAULTS[k]) for k in opts), "switch unknown to the loaded library"
    try:
        for k, v in opts.items():
            assert set_option(k, v)
        test_xattn_seeded_vs_oracle(3, 200, 2, 256, 192)
        test_xattn_seeded_vs_oracle(2, 128, 1, 768, 768)
        test_resampler_seeded_vs_oracle(4, 1, 50, 256, 2)
        test_resampler_seeded_vs_oracle(2, 2, 33, 128, 1)
        torch.cuda.synchronize()
    finally:
        for k in opts:
            set_option(k, OPTION_DEFAULTS[k])


def test_programmatic_dependent_launch_under_graph_capture():
    """FM_OPT_PDL inside a captured CUDA graph (how bench.py runs the step): replayed results == eager results."""
    from tests._gpu_util import set_option
    assert set_option("pdl", 1), "switch unknown to the loaded library"
    try:
        params = O.seeded_params(O.xattn_param_shapes(256, 192), 5)
        m = GatedCrossAttentionBlock(dim=256, dim_visual=192)
        m.load_state_dict(params); m = m.to(DEV)
        g = torch.Generator().manual_seed(2)
        y = torch.randn(2, 96, 256, generator=g).to(torch.bfloat16).to(DEV).requires_grad_(True)
        vis = torch.randn(2, 1, 64, 192, generator=g).to(torch.bfloat16).to(DEV).requires_grad_(True)
        ml = torch.zeros(2, 96, dtype=torch.long, device=DEV); ml[:, 1] = 1

        def step():
            m.zero_grad(set_to_none=True); y.grad = None; vis.grad = None
            out, _ = m(y, vis, ml)
            out.float().square().mean().backward()
            return out

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                ref = step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ref_out, ref_dy = ref.detach().clone(), y.grad.detach().clone()
        ref_gw = m.ffw[1].weight.grad.detach().clone()
        graph = torch.cuda.CUDAGraph()
        m.zero_grad(set_to_none=True); y.grad = None; vis.grad = None
        with torch.cuda.graph(graph):
            out = step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref_out)
        assert rel_err(y.grad, ref_dy) < 1e-5
        assert rel_err(m.ffw[1].weight.grad, ref_gw) < 1e-5
    finally:
        set_option("pdl", 1)          # back to the default


def test_deferred_side_join():
    """defer_join=1: fm_xattn_bwd returns with its weight-gradient GEMMs still on the side stream; gradients are complete after
    Fn.side_join() and equal to the joined run's; buffers are parked meanwhile; a second backward into existing .grad joins at once."""
    from flamingo_mini_b200 import functional as Fn
    from tests._gpu_util import set_option
    assert Fn.set_defer_join(False), "fm_side_join / defer_join not exported by the loaded library"
    params = O.seeded_params(O.xattn_param_shapes(256, 192), 5)
    m = GatedCrossAttentionBlock(dim=256, dim_visual=192)
    m.load_state_dict(params); m = m.to(DEV)
    g = torch.Generator().manual_seed(2)
    y = torch.randn(2, 96, 256, generator=g).to(torch.bfloat16).to(DEV).requires_grad_(True)
    vis = torch.randn(2, 1, 64, 192, generator=g).to(torch.bfloat16).to(DEV).requires_grad_(True)
    ml = torch.zeros(2, 96, dtype=torch.long, device=DEV); ml[:, 1] = 1
    junk = torch.randn(512, 512, device=DEV)

    def step(defer):
        assert set_option("defer_join", int(defer))
        m.zero_grad(set_to_none=True); y.grad = None; vis.grad = None
        out, _ = m(y, vis, ml)
        out.float().square().mean().backward()
        if defer:
            assert len(Fn._PENDING) == 1                  # saved / scratch / dy_out parked until the join
            for _ in range(4):
                junk @ junk                               # caller-side work the side stream overlaps with
            Fn.side_join()
        assert not Fn._PENDING
        if DEV != "cpu":
            torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters()}, y.grad.detach().clone()

    try:
        ref, ref_dy = step(False)
        got, got_dy = step(True)
        assert torch.equal(got_dy, ref_dy)
        for n in ref:                                     # same kernels; sums folded with atomics (gates, LayerNorm dgamma/dbeta
            assert rel_err(got[n], ref[n]) < 1e-4, n      # across row groups) may differ in the last bit
        # gradient accumulation (existing .grad): autograd adds into it right away, so the backward joins by itself
        assert set_option("defer_join", 1)
        out, _ = m(y, vis, ml)
        out.float().square().mean().backward()
        assert not Fn._PENDING
        assert rel_err(m.ffw[1].weight.grad, 2 * ref["ffw.1.weight"]) < 1e-3
    finally:
        Fn.set_defer_join(False)


# ---- stand-alone (inference) forwards of the sub-modules, against the oracle's restatement of the same reference functions
@pytest.mark.parametrize("act", ["gelu", "sqrelu", "relu"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_feed_forward_standalone(act, dtype):
    from flamingo_mini_b200 import FeedForward
    D = 256
    ff = FeedForward(D, mult=4, act=act).to(DEV)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        ff[0].weight.copy_(torch.randn(D, generator=g) * 0.2 + 1); ff[0].bias.copy_(torch.randn(D, generator=g) * 0.1)
    x = torch.randn(3, 50, D, generator=g).to(dtype)
    p = {"0.weight": ff[0].weight, "0.bias": ff[0].bias, "1.weight": ff[1].weight, "3.weight": ff[3].weight}
    ref = O.feed_forward(x.double(), {k: v.detach().double().cpu() for k, v in p.items()}, "", act)
    with torch.no_grad():
        out = ff(x.to(DEV))
    assert out.shape == x.shape and out.dtype == dtype
    _close(out, ref, 2e-2, "ffw")
    # with gradients (standalone._FeedForwardFn: the same primitives, backward = 4 GEMMs + LayerNorm backward) vs oracle autograd
    cot = torch.randn(3, 50, D, generator=g).to(dtype)
    p64 = {k: v.detach().double().cpu().requires_grad_(True) for k, v in p.items()}
    x64 = x.double().requires_grad_(True)
    O.feed_forward(x64, p64, "", act).backward(cot.double())
    xd = x.to(DEV).requires_grad_(True)
    out_g = ff(xd)
    assert torch.equal(out_g, out)                       # same forward kernels with and without the saved act'
    out_g.backward(cot.to(DEV))
    ew = act != "relu"          # ReLU's derivative is a step: pre-activations within bf16 rounding of 0 flip whole gradient rows
    _close(xd.grad, x64.grad, 5e-2, "ffw dx", elementwise=ew)
    assert xd.grad.dtype == dtype
    for name, par in (("0.weight", ff[0].weight), ("0.bias", ff[0].bias), ("1.weight", ff[1].weight), ("3.weight", ff[3].weight)):
        _close(par.grad, p64[name].grad, 5e-2, "ffw d" + name, elementwise=ew)


@pytest.mark.parametrize("heads", [8, 3])
def test_masked_cross_attention_standalone(heads):
    from flamingo_mini_b200 import _lib
    assert _lib.has("fm_xattn_core_fwd"), "entry point not exported by the loaded library"
    D, Dv, B, S, N = 256, 192, 2, 70, 3
    params = O.seeded_params(O.xattn_param_shapes(D, Dv, heads=heads), 11)
    blk = GatedCrossAttentionBlock(dim=D, dim_visual=Dv, heads=heads)
    blk.load_state_dict(params); blk = blk.to(DEV)
    g = torch.Generator().manual_seed(8)
    y = torch.randn(B, S, D, generator=g).to(torch.bfloat16)
    vis = torch.randn(B, N, 64, Dv, generator=g).to(torch.bfloat16)
    ml = torch.zeros(B, S, dtype=torch.long)
    ml[0, 3] = 1; ml[0, 30] = 1; ml[0, 50] = 1; ml[0, 60] = 1     # 4 tags, 3 images: the last rows see the uniform average
    ml[1, 10] = 1                                                   # rows 0..9 of sample 1 precede every image: exact zeros
    p64 = {k: v.double() for k, v in params.items()}
    ref, (rk, rv) = O.masked_cross_attention(y.double(), ml, vis.double(), p64, "attn.", heads=heads, output_kv=True)
    with torch.no_grad():
        out, (k, v) = blk.attn(y.to(DEV), ml.to(DEV), vis.to(DEV), output_kv=True)
        _close(out, ref, 2e-2, "attn out")
        assert not out[1, :10].any()
        _close(k, rk, 1e-2, "k"); _close(v, rv, 1e-2, "v")
        oc, none = blk.attn(y[:, -5:].to(DEV), ml.to(DEV), None, previous_kv=(k, v))     # cached decoding: last 5 tokens
        assert none is None
        _close(oc, ref[:, -5:], 2e-2, "cached")


@pytest.mark.parametrize("heads", [8, 3])
def test_perceiver_attention_standalone(heads):
    from flamingo_mini_b200 import _lib
    assert _lib.has("fm_resampler_core_fwd"), "entry point not exported by the loaded library"
    Dv, b, n1 = 128, 3, 77
    params = O.seeded_params(O.resampler_param_shapes(Dv, 1, heads=heads), 12)
    res = PerceiverResampler(dim=Dv, depth=1, heads=heads)
    res.load_state_dict(params); res = res.to(DEV)
    g = torch.Generator().manual_seed(6)
    feats = torch.randn(b, n1, Dv, generator=g).to(torch.bfloat16)
    lat = torch.randn(b, 64, Dv, generator=g)
    ref = O.perceiver_attention(feats.double(), lat.double(), {k: v.double() for k, v in params.items()}, "layers.0.0.", heads=heads)
    with torch.no_grad():
        out = res.layers[0][0](feats.to(DEV), lat.to(DEV))
    assert out.shape == (b, 64, Dv) and out.dtype == torch.float32
    _close(out, ref, 2e-2, "perceiver attention")


@pytest.mark.parametrize("heads", [8, 2])
def test_standalone_attention_modules_with_gradients(heads):
    """MaskedCrossAttention / PerceiverAttentionLayer called on their own under autograd (standalone.py: primitives + the core
    backward entry points of the C ABI) against the oracle's autograd on the same reference functions."""
    from flamingo_mini_b200 import _lib
    assert _lib.has("fm_xattn_core_bwd"), "entry point not exported by the loaded library"
    # ---- MaskedCrossAttention (gated_cross_attention.py:42-131)
    D, Dv, B, S, N = 128, 192, 2, 70, 2
    params = O.seeded_params(O.xattn_param_shapes(D, Dv, heads=heads), 21)
    blk = GatedCrossAttentionBlock(dim=D, dim_visual=Dv, heads=heads)
    blk.load_state_dict(params); blk = blk.to(DEV)
    g = torch.Generator().manual_seed(3)
    y = torch.randn(B, S, D, generator=g).to(torch.bfloat16)
    vis = torch.randn(B, N, 64, Dv, generator=g).to(torch.bfloat16)
    ml = torch.zeros(B, S, dtype=torch.long); ml[:, 2] = 1; ml[:, 40] = 1; ml[1, 60] = 1      # sample 1: a third tag, only 2 images
    cot = torch.randn(B, S, D, generator=g).to(torch.bfloat16)
    fn = lambda i, p: O.masked_cross_attention(i[0], i[2], i[1], p, "attn.", heads=heads)[0]            # noqa: E731
    o_out, o_gin, o_gp = _oracle_grads(fn, [y, vis, ml], params, cot)
    yd, vd = y.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    out, kv = blk.attn(yd, ml.to(DEV), vd)
    assert kv is None
    _close(out, o_out, 2e-2, "attn out")
    out.backward(cot.to(DEV))
    _close(yd.grad, o_gin[0], 5e-2, "attn dy")
    _close(vd.grad, o_gin[1], 6e-2, "attn dvis")
    for n, p_ in blk.attn.named_parameters():
        _close(p_.grad, o_gp["attn." + n], 6e-2, "attn d" + n)
    # ---- PerceiverAttentionLayer (perceiver_resampler.py:32-96)
    Dr, b, n1 = 128, 3, 77
    rparams = O.seeded_params(O.resampler_param_shapes(Dr, 1, heads=heads), 22)
    res = PerceiverResampler(dim=Dr, depth=1, heads=heads)
    res.load_state_dict(rparams); res = res.to(DEV)
    layer = res.layers[0][0]
    feats = torch.randn(b, n1, Dr, generator=g).to(torch.bfloat16)
    lat = torch.randn(b, 64, Dr, generator=g)
    cot2 = torch.randn(b, 64, Dr, generator=g)
    fn2 = lambda i, p: O.perceiver_attention(i[0], i[1], p, "layers.0.0.", heads=heads)               # noqa: E731
    r_out, r_gin, r_gp = _oracle_grads(fn2, [feats, lat], rparams, cot2)
    fd, ld = feats.to(DEV).requires_grad_(True), lat.to(DEV).requires_grad_(True)
    out2 = layer(fd, ld)
    _close(out2, r_out, 2e-2, "perceiver out")
    out2.backward(cot2.to(DEV))
    _close(fd.grad, r_gin[0], 6e-2, "perceiver dfeatures")
    _close(ld.grad, r_gin[1], 6e-2, "perceiver dlatents")
    for n, p_ in layer.named_parameters():
        _close(p_.grad, r_gp["layers.0.0." + n], 6e-2, "perceiver d" + n)
