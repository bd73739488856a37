"""Generate golden vectors by running the UNMODIFIED reference modules.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [--only res_h4,xattn_h2]      # --only: write just these cases (others stay untouched)

Imports ``/root/reference/flamingo_mini/{perceiver_resampler,gated_cross_attention}.py``
with a tests-only two-function shim for the missing ``einops_exts`` package, loads
deterministic parameters (``oracle.flamingo_oracle.seeded_params``) into the reference
modules, runs forward + autograd backward in fp64 on seeded inputs and writes
``tests/golden/*.pt``.  Large parameter gradients are stored as digests
(sum, L2 norm, two seeded random projections) to keep the fixtures small.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("FLAMINGO_REF", "/root/reference")


def _install_shim():
    from einops import rearrange, repeat
    m = types.ModuleType("einops_exts")
    m.rearrange_many = lambda ts, pat, **kw: tuple(rearrange(t, pat, **kw) for t in ts)
    m.repeat_many = lambda ts, pat, **kw: tuple(repeat(t, pat, **kw) for t in ts)
    sys.modules["einops_exts"] = m


def _import_reference():
    """Import the two hot-path files as a synthetic package (avoids flamingo_mini/__init__,
    which pulls transformers/PIL-heavy modules we do not need)."""
    import importlib.util
    pkg = types.ModuleType("flamingo_mini_ref")
    pkg.__path__ = [os.path.join(REF, "flamingo_mini")]
    sys.modules["flamingo_mini_ref"] = pkg
    mods = {}
    for name in ("utils", "perceiver_resampler", "gated_cross_attention"):
        spec = importlib.util.spec_from_file_location(
            f"flamingo_mini_ref.{name}", os.path.join(REF, "flamingo_mini", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"flamingo_mini_ref.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def digest(name: str, g: torch.Tensor) -> torch.Tensor:
    """Compact, order-sensitive summary of a big gradient tensor."""
    gen = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    r = torch.randn(2, g.numel(), generator=gen, dtype=torch.float64)
    gf = g.reshape(-1).double()
    return torch.stack([gf.sum(), gf.norm(), r[0] @ gf, r[1] @ gf])


def hash_name(name: str) -> int:
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) % (2 ** 61)
    return h


def pack_grads(named):
    out = {}
    for n, g in named.items():
        out[n] = {"full": g.float()} if g.numel() <= 8192 else {"digest": digest(n, g)}
    return out


def main():
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    _install_shim()
    ref = _import_reference()
    from oracle.flamingo_oracle import resampler_param_shapes, seeded_params, xattn_param_shapes

    torch.manual_seed(0)
    dt = torch.float64

    # ------------------------------------------------------------------ resampler cases
    res_cases = {
        "res_img": dict(dim=128, depth=2, b=3, T=None, f=10, act="gelu", seed=11),
        "res_vid": dict(dim=64, depth=1, b=2, T=3, f=7, act="sqrelu", seed=12),
        "res_relu": dict(dim=64, depth=1, b=1, T=None, f=5, act="relu", seed=13),
        # non-default constructor arguments (perceiver_resampler.py:100-111): 4 heads, ff_mult 2, two frames
        "res_h4": dict(dim=128, depth=1, b=2, T=2, f=9, act="gelu", seed=14, heads=4, ff_mult=2),
    }
    for name, c in res_cases.items():
        if only is not None and name not in only:
            continue
        heads, ff_mult = c.get("heads", 8), c.get("ff_mult", 4)
        shapes = resampler_param_shapes(c["dim"], c["depth"], heads=heads, ff_mult=ff_mult)
        params = seeded_params(shapes, c["seed"], dtype=dt)
        m = ref["perceiver_resampler"].PerceiverResampler(dim=c["dim"], depth=c["depth"], heads=heads, ff_mult=ff_mult,
                                                          act=c["act"]).to(dt)
        missing = m.load_state_dict(params, strict=True)
        g = torch.Generator().manual_seed(100 + c["seed"])
        shape = (c["b"], c["f"], c["dim"]) if c["T"] is None else (c["b"], c["T"], c["f"], c["dim"])
        x = torch.randn(shape, generator=g, dtype=torch.float32).to(dt).requires_grad_(True)   # fp32-exact inputs
        w = torch.randn((c["b"], 64, c["dim"]), generator=g, dtype=torch.float32).to(dt)      # cotangent
        out = m(x)
        (out * w).sum().backward()
        fx = dict(case=c, x=x.detach().float(), cot=w.float(), out=out.detach().float(),
                  dx=x.grad.float(), dparams=pack_grads({n: p.grad for n, p in m.named_parameters()}),
                  n_params=sum(p.numel() for p in m.parameters()))
        torch.save(fx, os.path.join(HERE, f"{name}.pt"))
        print(name, "out", tuple(out.shape), "params", fx["n_params"], missing)

    # T > num_time_embeds must raise (perceiver_resampler.py:166)
    m = ref["perceiver_resampler"].PerceiverResampler(dim=64, depth=1)
    try:
        m(torch.randn(1, 5, 3, 64))
        raised = False
    except RuntimeError:
        raised = True
    assert raised

    # ------------------------------------------------------------------ gated xattn cases
    def media(b, s, marks):
        ml = torch.zeros(b, s, dtype=torch.int64)
        for r, cols in enumerate(marks):
            for cidx in cols:
                ml[r, cidx] = 1
        return ml

    x_cases = {
        # row0: text before first image (tt==0) then two images; row1: 3 markers but only 2 media -> uniform rows
        "xattn_edge": dict(dim=192, dim_visual=128, b=2, s=24, n=2, act="gelu", seed=21,
                           marks=[[5, 14], [0, 8, 17]]),
        "xattn_sq": dict(dim=64, dim_visual=64, b=3, s=9, n=1, act="sqrelu", seed=22,
                         marks=[[0], [3], []]),
        # non-default constructor arguments (gated_cross_attention.py:136-146): 2 heads, ff_mult 2; three images, the second row
        # only ever reaches the second one
        "xattn_h2": dict(dim=128, dim_visual=64, b=2, s=20, n=3, act="gelu", seed=23, heads=2, ff_mult=2,
                         marks=[[0, 6, 13], [2, 9]]),
    }
    for name, c in x_cases.items():
        if only is not None and name not in only:
            continue
        heads, ff_mult = c.get("heads", 8), c.get("ff_mult", 4)
        shapes = xattn_param_shapes(c["dim"], c["dim_visual"], heads=heads, ff_mult=ff_mult)
        params = seeded_params(shapes, c["seed"], dtype=dt)
        m = ref["gated_cross_attention"].GatedCrossAttentionBlock(
            dim=c["dim"], dim_visual=c["dim_visual"], heads=heads, ff_mult=ff_mult, act=c["act"]).to(dt)
        m.load_state_dict(params, strict=True)
        g = torch.Generator().manual_seed(100 + c["seed"])
        y = torch.randn((c["b"], c["s"], c["dim"]), generator=g, dtype=torch.float32).to(dt).requires_grad_(True)
        vis = torch.randn((c["b"], c["n"], 64, c["dim_visual"]), generator=g, dtype=torch.float32).to(dt).requires_grad_(True)
        w = torch.randn((c["b"], c["s"], c["dim"]), generator=g, dtype=torch.float32).to(dt)
        ml = media(c["b"], c["s"], c["marks"])
        out, kv = m(y, vis, ml, previous_kv=None, output_kv=True)
        (out * w).sum().backward()
        fx = dict(case=c, y=y.detach().float(), vis=vis.detach().float(), media_locations=ml, cot=w.float(),
                  out=out.detach().float(), k=kv[0].detach().float(), v=kv[1].detach().float(),
                  dy=y.grad.float(), dvis=vis.grad.float(),
                  dparams=pack_grads({n: p.grad for n, p in m.named_parameters()}),
                  n_params=sum(p.numel() for p in m.parameters()))
        # cached decode step: last 3 tokens with previous_kv (gated_cross_attention.py:88-104)
        with torch.no_grad():
            out_c, _ = m(y[:, -3:].detach(), None if False else vis.detach(), ml, previous_kv=(kv[0].detach(), kv[1].detach()))
        fx["out_cached_last3"] = out_c.float()
        torch.save(fx, os.path.join(HERE, f"{name}.pt"))
        print(name, "out", tuple(out.shape), "params", fx["n_params"])

    # gate = 0 -> identity (torch.equal) fact, recorded as a flag
    m = ref["gated_cross_attention"].GatedCrossAttentionBlock(dim=64, dim_visual=64)
    y = torch.randn(2, 5, 64)
    o, _ = m(y, torch.randn(2, 1, 64, 64), media(2, 5, [[0], [1]]))
    assert torch.equal(o, y)

    # known-answer parameter counts (examples/model_stats.ipynb:1583-1584,1605)
    n_res = sum(p.numel() for p in ref["perceiver_resampler"].PerceiverResampler(dim=1024, depth=6).parameters())
    n_x = sum(p.numel() for p in ref["gated_cross_attention"].GatedCrossAttentionBlock(dim=768, dim_visual=1024).parameters())
    assert n_res == 63023104 and n_x == 6556674, (n_res, n_x)
    print("param-count KATs ok", n_res, n_x)


if __name__ == "__main__":
    main()
