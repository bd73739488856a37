#!/usr/bin/env python
"""Cached generation through the CUDA gated-xattn blocks (modeling_flamingo.py:464-523,550-605 and
gated_cross_attention.py:88-104,128-131 of the reference): prefix forward with use_cache, then single-token decode steps that
reuse the cached keys/values (`previous_kv`), timed per token.

    python tools/decode_bench.py [--workload c2] [--batch 8] [--prefix 32] [--tokens 64]

Prints one JSON line: ms per decode step for the whole model (frozen HF LM + 12 gated xattn blocks), the library's share
(per-kernel events), launches per step, and a check that cached greedy decoding reproduces the un-cached argmax.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--prefix", type=int, default=32)
    ap.add_argument("--tokens", type=int, default=64)
    args = ap.parse_args()
    from flamingo_mini_b200 import _lib
    lib = _lib.load()
    w = dict(bench.WORKLOADS[args.workload])
    dev = torch.device("cuda", 0)
    model = bench.build_model(w, dev).eval()
    B, P = args.batch, args.prefix
    g = torch.Generator().manual_seed(5)
    clip = torch.randn(B * w["N"], 1, w["F"], w["Dv"], generator=g).to(torch.bfloat16).to(dev)
    ids = torch.randint(0, w["lm_config"]["vocab_size"], (B, P), generator=g).to(dev)
    ml = torch.zeros(B, P, dtype=torch.int64, device=dev)
    ml[:, 0] = 1
    with torch.no_grad():
        vf = model.flamingo.resampler(clip).reshape(B, w["N"], 64, w["Dv"])

        def decode(n, profile=False):
            cur, cml = ids, ml
            out = model(input_ids=cur, media_locations=cml, visual_features=vf, use_cache=True, attention_mask=torch.ones_like(cur))
            past = out.past_key_values
            nxt = out.logits[:, -1].argmax(-1, keepdim=True)
            toks = [nxt]
            torch.cuda.synchronize()
            l0 = lib.fm_launch_count()
            if profile:
                lib.fm_profile_enable(1)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(n):
                cur = torch.cat([cur, nxt], 1)
                cml = torch.cat([cml, torch.zeros_like(cml[:, :1])], 1)
                out = model(input_ids=nxt, media_locations=cml, past_key_values=past, use_cache=True, attention_mask=torch.ones_like(cur))
                past = out.past_key_values
                nxt = out.logits[:, -1].argmax(-1, keepdim=True)
                toks.append(nxt)
            ev1.record()
            torch.cuda.synchronize()
            prof = bench.parse_profile(lib) if profile else None
            if profile:
                lib.fm_profile_enable(0)
            return ev0.elapsed_time(ev1) / n, (lib.fm_launch_count() - l0) / n, torch.cat(toks, 1), prof

        decode(4)                                   # warm-up
        ms, launches, toks, _ = decode(args.tokens)
        _, _, _, prof = decode(8, profile=True)
        lib_ms = sum(v["ms"] for v in prof.values()) / 8
        # un-cached check of the first few generated tokens
        cur, cml = ids, ml
        agree, total = 0, 0
        for t in range(min(6, args.tokens)):
            lg = model(input_ids=cur, media_locations=cml, visual_features=vf, attention_mask=torch.ones_like(cur)).logits[:, -1].float()
            top2 = lg.topk(2).values
            sure = (top2[:, 0] - top2[:, 1]) > 5e-2 * lg.abs().amax(-1)
            agree += int((lg.argmax(-1)[sure] == toks[sure, t]).sum())
            total += int(sure.sum())
            cur = torch.cat([cur, toks[:, t:t + 1]], 1)
            cml = torch.cat([cml, torch.zeros_like(cml[:, :1])], 1)
    n_blocks = len(list(model.flamingo.get_modified_layers()))
    print(json.dumps({"metric": "cached decode step (whole model, eager launches)", "workload": args.workload, "batch": B, "prefix": P,
                      "ms_per_token": ms, "tokens_per_s": B / ms * 1e3, "library_launches_per_token": launches,
                      "library_kernel_ms_per_token": lib_ms, "xattn_blocks": n_blocks,
                      "library_us_per_block_per_token": lib_ms / n_blocks * 1e3,
                      "cached_equals_uncached_argmax": f"{agree}/{total} unambiguous positions"}))


if __name__ == "__main__":
    main()
