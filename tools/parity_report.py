#!/usr/bin/env python
"""Logits / loss / gradient parity of the CUDA hot path at MODEL level, on the GPU, written as a markdown table
(profiles/r02_parity.md) — VERDICT r1 "Next round" 1(d) — plus the practical same-box comparison point BASELINE.md §4 names:
the reference modules in eager PyTorch (cuBLAS + ATen) on the B200, fp32 and under bf16 autocast.

    python tools/parity_report.py [--out profiles/r02_parity.md] [--workload c2] [--batch 8] [--time-steps 10]

Three models share ONE set of weights (random init, seed 0, gates opened to 0.5 like bench.py):
  ours        FlamingoModel with the sm_100a PerceiverResampler / GatedCrossAttentionBlocks (bf16 tensor-core operands)
  ref-fp32    the oracle restatement of the reference modules (oracle/flamingo_oracle.py, plain torch) on the GPU in fp32
              (TF32 off) — "the reference forward on identical inputs"
  ref-bf16    the same oracle modules under torch.autocast(bfloat16) — what the reference's own mixed-precision recipe
              (training/train.sh:24, fp16 AMP) computes; its distance from ref-fp32 is the error class of 16-bit operands
The frozen HF LM runs in fp32 in all three (so only the hot path differs).  Errors are reported against ref-fp32.
The oracle is used here as the CHECKER only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def build_pair(w, dev, lm_dtype=torch.float32):
    from flamingo_mini_b200.configuration_flamingo import FlamingoConfig
    from flamingo_mini_b200.modeling_flamingo import FlamingoModel
    from oracle.oracle_modules import swap_in_oracle
    cfg_kw = dict(lm=w["lm"], dim=w["D"], dim_visual=w["Dv"], xattn_every=w["xattn_every"],
                  lm_config=dict(w["lm_config"], **({"resid_pdrop": 0.0, "embd_pdrop": 0.0, "attn_pdrop": 0.0} if w["lm"].startswith("gpt") else
                                                    {"dropout": 0.0, "attention_dropout": 0.0, "activation_dropout": 0.0})),
                  clip_config=bench.CLIP_TINY)
    torch.manual_seed(0)
    ours = FlamingoModel(FlamingoConfig(**cfg_kw))
    with torch.no_grad():
        for layer in ours.flamingo.get_modified_layers():
            layer.xattn_block.alpha_attn.fill_(0.5)
            layer.xattn_block.alpha_ffw.fill_(0.5)
    sd = {k: v.clone() for k, v in ours.state_dict().items()}
    torch.manual_seed(0)
    ref = FlamingoModel(FlamingoConfig(**cfg_kw))
    ref.load_state_dict(sd)
    swap_in_oracle(ref, copy_weights=True)
    ours.flamingo.lm.to(lm_dtype); ours.flamingo.lm_head.to(lm_dtype)
    return ours.to(dev).eval(), ref.float().to(dev).eval()


def run(model, w, clip, ids, ml, autocast=False, backward=True):
    model.zero_grad(set_to_none=True)
    B = ids.shape[0]
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with ctx:
        vf = model.flamingo.resampler(clip).reshape(B, w["N"], 64, w["Dv"])
        out = model(input_ids=ids, media_locations=ml, visual_features=vf, labels=ids, attention_mask=torch.ones_like(ids))
    grads = {}
    if backward:
        out.loss.float().backward()
        grads = {n: p.grad.detach().float() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}
    return out.logits.detach().float(), float(out.loss.detach()), grads


def stats(got, ref):
    d = (got - ref).abs()
    rms = ref.square().mean().sqrt().item()
    rel = (d.norm() / (ref.norm() + 1e-30)).item()
    within = lambda rtol, atol: (d <= atol + rtol * ref.abs()).float().mean().item()      # noqa: E731
    return dict(rel_l2=rel, max_abs=d.max().item(), rms_ref=rms, within_rtol1e3_atol1e3rms=within(1e-3, 1e-3 * rms),
                within_rtol1e2_atol1e2rms=within(1e-2, 1e-2 * rms),
                rtol_needed_atol_1e3rms=((d - 1e-3 * rms).clamp(min=0) / ref.abs().clamp(min=1e-12)).quantile(0.999).item()
                if d.numel() < 16_000_000 else float("nan"))


def grad_table(got, ref, ref_names_map):
    rows = []
    for n, g in got.items():
        rn = ref_names_map(n)
        if rn not in ref:
            continue
        r = ref[rn]
        if r.norm().item() < 1e-12:
            continue
        rows.append((n, ((g - r.reshape(g.shape)).norm() / r.norm()).item()))
    return rows


def oracle_name(n: str) -> str:
    """our parameter name -> the oracle module's flattened name ('a.b' -> 'a__b' below the swapped module)"""
    for mark in (".xattn_block.", "flamingo.resampler."):
        if mark in n:
            head, tail = n.split(mark, 1)
            return head + mark + tail.replace(".", "__")
    return n


def time_eager(model, w, clip, ids, ml, autocast, steps):
    for _ in range(2):
        run(model, w, clip, ids, ml, autocast)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        run(model, w, clip, ids, ml, autocast)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity.md"))
    ap.add_argument("--workloads", default="tiny,c2")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--time-steps", type=int, default=10)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    lines = ["# r02 model-level parity on the B200 (tools/parity_report.py)", "",
             "Reference = oracle restatement of the reference modules in fp32 on the same GPU (TF32 off), same weights, same inputs; frozen HF LM in fp32 in every arm.",
             "`ours` = CUDA hot path (bf16 tensor-core operands, fp32 accumulation/statistics).  `ref-bf16` = the reference modules under bf16 autocast.",
             "", "| workload | arm | logits rel L2 | logits max abs | rms(ref logits) | frac within rtol 1e-3 (+1e-3 rms) | frac within rtol 1e-2 (+1e-2 rms) | loss (ref) | loss (arm) | worst grad rel L2 | median grad rel L2 |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    records = []
    for name in args.workloads.split(","):
        w = dict(bench.WORKLOADS[name])
        B = min(args.batch, w["B"])
        ours, ref = build_pair(w, dev)
        clip, ids, ml = bench.make_batch(w, B, dev, 4321, torch.float32)
        lg_ref, loss_ref, g_ref = run(ref, w, clip, ids, ml)
        lg_bf, loss_bf, g_bf = run(ref, w, clip, ids, ml, autocast=True)
        lg_our, loss_our, g_our = run(ours, w, clip.to(torch.bfloat16), ids, ml)
        for arm, lg, loss, g, mapper in (("ours", lg_our, loss_our, g_our, oracle_name), ("ref-bf16", lg_bf, loss_bf, g_bf, lambda n: n)):
            st = stats(lg, lg_ref)
            gt = grad_table(g, g_ref, mapper)
            gerrs = sorted(e for _, e in gt)
            worst = max(gt, key=lambda t: t[1]) if gt else ("-", float("nan"))
            lines.append(f"| {name} (B={B}) | {arm} | {st['rel_l2']:.2e} | {st['max_abs']:.2e} | {st['rms_ref']:.2e} | {st['within_rtol1e3_atol1e3rms']:.4f} | "
                         f"{st['within_rtol1e2_atol1e2rms']:.4f} | {loss_ref:.5f} | {loss:.5f} | {worst[1]:.2e} ({worst[0].split('flamingo.')[-1]}) | "
                         f"{(gerrs[len(gerrs) // 2] if gerrs else float('nan')):.2e} |")
            records.append(dict(workload=name, batch=B, arm=arm, logits=st, loss_ref=loss_ref, loss=loss, worst_grad=worst, n_grads=len(gt)))
        del ours, ref
        torch.cuda.empty_cache()
    # ---- the practical same-box comparison: reference modules in eager PyTorch on the B200
    lines += ["", "## Reference modules in eager PyTorch on the same B200 (cuBLAS + ATen), C2 at the benchmark batch", "",
              "fwd+bwd of the whole step (resampler + frozen gpt2 with 12 gated xattn blocks + loss), eager launches, no CUDA graph; "
              "the frozen LM runs bf16 (as in bench.py) for the bf16 arms.", "",
              "| arm | ms/step | samples/s |", "|---|---|---|"]
    w = dict(bench.WORKLOADS["c2"])
    B = w["B"]
    for arm, lm_dtype, autocast in (("reference modules, fp32 (TF32 off)", torch.float32, False),
                                    ("reference modules, bf16 autocast, bf16 LM", torch.bfloat16, True)):
        ours, ref = build_pair(w, dev)
        del ours
        ref.train()
        if lm_dtype != torch.float32:
            ref.flamingo.lm.to(lm_dtype); ref.flamingo.lm_head.to(lm_dtype)
        clip, ids, ml = bench.make_batch(w, B, dev, 1234, torch.float32 if not autocast else torch.bfloat16)
        ms = time_eager(ref, w, clip, ids, ml, autocast, args.time_steps)
        lines.append(f"| {arm} | {ms:.2f} | {B / ms * 1e3:.0f} |")
        records.append(dict(timing=arm, ms_per_step=ms, samples_per_s=B / ms * 1e3))
        del ref
        torch.cuda.empty_cache()
    ours, ref = build_pair(w, dev, lm_dtype=torch.bfloat16)
    del ref
    ours.train()
    clip, ids, ml = bench.make_batch(w, B, dev, 1234, torch.bfloat16)
    t0 = time.time()
    ms = time_eager(ours, w, clip, ids, ml, False, args.time_steps)
    lines.append(f"| this library, eager launches (no CUDA graph; bench.py replays a graph) | {ms:.2f} | {B / ms * 1e3:.0f} |")
    records.append(dict(timing="ours eager", ms_per_step=ms, samples_per_s=B / ms * 1e3, wall_s=time.time() - t0))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")
    open(args.out.replace(".md", ".json"), "w").write(json.dumps(records, indent=1, default=str))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
