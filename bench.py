#!/usr/bin/env python
"""bench.py — image-text samples/sec (fwd+bwd) of the flamingo-mini hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5|tiny]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the training hot path over one synthetic batch: PerceiverResampler(random CLIP patch
features) -> frozen HF language model with a GatedCrossAttentionBlock spliced before every xattn_every-th layer
(random token ids, labels = ids) -> loss -> backward -> data-parallel all-reduce (mean) of every trainable gradient.
The frozen CLIP tower is bypassed (north_star: inputs are random CLIP patch features); the frozen LM stays stock
PyTorch/HF in bf16.  No optimizer step: the metric is "samples/sec (fwd+bwd)" (BASELINE.json).

Output: ONE JSON line on rank 0 (contract in the task statement): value = whole-job samples/s with inputs resident
in HBM; e2e = same through the public module API with pinned-host inputs (H2D + loss D2H inside the timed region);
roofline = the tcgen05 GEMM template (all instantiations, per-instantiation table) from device-side kernel durations:
right after the timed region the same step is captured once more with the library's side stream off and replayed
under CUPTI activity tracing, its records attributed through the library's launch log (fm_profile_log); xattn = the
"xattn TFLOPS vs peak" key of BASELINE.json over every kernel fm_xattn_fwd/bwd launch; cpu_baseline = the CPU oracle
port on the host cores.

--impl reference: the reference's own CPU path for the same step.  /root/reference (pure Python) cannot travel to
the GPU box, so this arm runs the oracle port (oracle/flamingo_oracle.py + oracle/oracle_model.py, pinned against the
reference by tests/golden) + the same stock HF LM in fp32 on all host cores at the workload's full per-GPU batch, the
number of timed steps bounded by --reference-max-seconds; it never imports flamingo_mini_b200.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (lm family, LM dims, CLIP tokens F, Dv, images N, seq S, batch per GPU, xattn_every)   
WORKLOADS = {
    "c2": dict(desc="gpt2 + ViT-B/32 bf16, 64 latents, xattn_every=1, seq 128, batch 32/GPU", lm="gpt2",
               lm_config=dict(n_embd=768, n_layer=12, n_head=12, vocab_size=50257, n_positions=1024),
               D=768, Dv=768, F=50, N=1, S=128, B=32, xattn_every=1),
    "c3": dict(desc="gpt2-large + ViT-L/14 bf16, resampler_depth=6, seq 256, batch 16/GPU", lm="gpt2-large",
               lm_config=dict(n_embd=1280, n_layer=36, n_head=20, vocab_size=50257, n_positions=1024),
               D=1280, Dv=1024, F=257, N=1, S=256, B=16, xattn_every=1),
    "c4": dict(desc="opt-1.3b + ViT-L/14, 4 interleaved images, seq 512, batch 8/GPU", lm="facebook/opt-1.3b",
               lm_config=dict(hidden_size=2048, num_hidden_layers=24, num_attention_heads=32, ffn_dim=8192,
                              vocab_size=50272, max_position_embeddings=2048, word_embed_proj_dim=2048),
               D=2048, Dv=1024, F=257, N=4, S=512, B=8, xattn_every=1),
    "c5": dict(desc="opt-6.7b + ViT-L/14, xattn_every=4, seq 1024, batch 4/GPU", lm="facebook/opt-6.7b",
               lm_config=dict(hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, ffn_dim=16384,
                              vocab_size=50272, max_position_embeddings=2048, word_embed_proj_dim=4096),
               D=4096, Dv=1024, F=257, N=1, S=1024, B=4, xattn_every=4),
    "tiny": dict(desc="tiny smoke configuration (not a benchmark)", lm="gpt2",
                 lm_config=dict(n_embd=128, n_layer=2, n_head=2, vocab_size=512, n_positions=128),
                 D=128, Dv=128, F=10, N=1, S=32, B=2, xattn_every=1),
}
CLIP_TINY = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, image_size=32,
                 patch_size=16)   # the CLIP tower is bypassed by the benchmark; keep it negligible


def hot_path_flops_per_sample(w, depth=6):
    """Algorithmic FLOPs (2*M*N*K per GEMM) of resampler + all xattn blocks, fwd+bwd, per sample."""
    I, Q = 512, 64
    Dv, D, F, S, N = w["Dv"], w["D"], w["F"], w["S"], w["N"]
    K = F + Q
    to_q, to_kv_lat, to_kv_med = 2 * Q * Dv * I, 4 * Q * Dv * I, 4 * F * Dv * I
    core, to_out, ffw = 4 * Q * K * I, 2 * Q * I * Dv, 16 * Q * Dv * Dv
    fwd_r = depth * (to_q + to_kv_lat + to_kv_med + core + to_out + ffw)
    bwd_r = depth * (2 * to_q + 2 * to_kv_lat + 1 * to_kv_med + 2 * core + 2 * to_out + 2 * ffw)
    n_blocks = len(range(0, w["lm_config"].get("n_layer", w["lm_config"].get("num_hidden_layers")), w["xattn_every"]))
    fwd_x = 2 * S * D * I + 4 * (N * Q) * Dv * I + 4 * S * 64 * I + 2 * S * I * D + 16 * S * D * D
    return N * (fwd_r + bwd_r) + n_blocks * 3 * fwd_x


def build_model(w, device, fused_loss=True):
    from flamingo_mini_b200.configuration_flamingo import FlamingoConfig
    from flamingo_mini_b200.modeling_flamingo import FlamingoModel
    torch.manual_seed(0)
    cfg = FlamingoConfig(lm=w["lm"], dim=w["D"], dim_visual=w["Dv"], xattn_every=w["xattn_every"],
                         lm_config=w["lm_config"], clip_config=CLIP_TINY, fused_cross_entropy=bool(fused_loss))
    model = FlamingoModel(cfg)
    with torch.no_grad():    # at the reference's init (alpha = 0) every block is the identity: open the gates
        for layer in model.flamingo.get_modified_layers():
            layer.xattn_block.alpha_attn.fill_(0.5)
            layer.xattn_block.alpha_ffw.fill_(0.5)
    model.flamingo.lm.to(torch.bfloat16)
    model.flamingo.lm_head.to(torch.bfloat16)
    return model.to(device)


def make_batch(w, B, device, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    clip = torch.randn(B * w["N"], 1, w["F"], w["Dv"], generator=g).to(dtype)
    vocab = w["lm_config"]["vocab_size"]
    ids = torch.randint(0, vocab, (B, w["S"]), generator=g)
    ml = torch.zeros(B, w["S"], dtype=torch.int64)
    for k in range(w["N"]):
        ml[:, (k * w["S"]) // w["N"]] = 1
    return clip.to(device), ids.to(device), ml.to(device)


def train_step(model, w, clip, ids, ml, reducer=None):
    B = ids.shape[0]
    vf = model.flamingo.resampler(clip).reshape(B, w["N"], 64, w["Dv"])
    out = model(input_ids=ids, media_locations=ml, visual_features=vf, labels=ids,
                attention_mask=torch.ones_like(ids))
    out.loss.backward()
    from flamingo_mini_b200 import functional as Fn
    if Fn._PENDING:                      # only with FM_B200_OPTS=defer_join=1: also lets a graph capture end
        Fn.side_join()
    if reducer is not None:
        reducer.finish()
    return out.loss


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin=0.0, t_end=float("inf")):
        """Summarise the samples that arrived inside [t_begin, t_end] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if t_begin <= t <= t_end + 0.06] or [r for _, r in self.rows[-3:]]
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from `ncu --set full` captures taken IN SITU (one launch
# each inside a C2 training step, cold caches as in the real step): profiles/r02_call2/ncu_in_situ_metrics.txt (tools/ncu_top_kernels.sh).
# Keyed by the library's profiler tag; only meaningful for the C2 workload.
NCU_TRAFFIC_BYTES = {
    "gemm_a0b0_epi1_bn256": 11.06e6 + 6.37e6,    # FFW1 4096x3072x768, ACT epilogue: operands read once; outputs mostly still in L2
    "gemm_a0b1_epi3_bn256": 36.24e6 + 1.12e6,    # DACT 4096x3072x768: operands + the saved act' (25 MB) read once
    "gemm_a1b1_epi0_bn128": 36.22e6 + 0.41e6,    # dW   3072x768x4096, fp32 output
    "gemm_a0b0_epi2_bn192": 11.31e6 + 0.00e6,    # to_out RESID 4096x768x512 (the first bn192 RESID launch of a block)
}


def merge_scopes(prof):
    """The library prefixes profiler tags with the module entry point they were launched from ("x/" gated xattn block, "r/"
    resampler); the per-kernel views merge them back."""
    out = {}
    for k, v in prof.items():
        base = k.split("/", 1)[1] if "/" in k else k
        o = out.setdefault(base, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for f in o:
            o[f] += v[f]
    return out


def scope_totals(prof, nprof, scope):
    rows = [v for k, v in prof.items() if k.startswith(scope + "/")]
    return {"ms_per_step": sum(v["ms"] for v in rows) / nprof, "launches_per_step": sum(v["launches"] for v in rows) // nprof,
            "launched_gflop_per_step": sum(v["flops"] for v in rows) / nprof / 1e9}


def make_roofline(prof, nprof, t_ms, peaks_path=None, timing=None):
    """prof: tag -> {launches, ms, flops, bytes} summed over `nprof` profiled steps with per-kernel CUDA events.
    The dominant kernel of the path is the tcgen05 GEMM template `gemm_tc_kernel<BN, A_MN, B_MN, EPI>` (one source
    kernel, ~2/3 of the library's GPU time); `achieved` is its algorithmic FLOPs per launch over its average launch
    duration across ALL its launches of the step, and `instantiations` breaks that down."""
    prof = merge_scopes(prof)
    peaks = {}
    try:
        peaks = json.load(open(peaks_path or os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (of measured; kernels are timed inside a long step)" if peaks
                else "fallback 1.4 PFLOP/s sustained (of fallback)")
    total_ms = sum(v["ms"] for v in prof.values()) or float("nan")
    kernels = {k: {"launches": v["launches"] // nprof, "ms_per_step": v["ms"] / nprof,
                   "tflops": (v["flops"] / v["ms"] / 1e9) if v["flops"] and v["ms"] else None,
                   "gbs": (v["bytes"] / v["ms"] / 1e6) if v["ms"] else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    gemms = {k: v for k, v in prof.items() if k.startswith("gemm_") and v["ms"] > 0}
    if not gemms:
        return None, kernels
    all_f = sum(v["flops"] for v in gemms.values())
    all_ms = sum(v["ms"] for v in gemms.values())
    all_n = sum(v["launches"] for v in gemms.values())
    ach = all_f / all_ms / 1e9
    inst = []
    for k, v in sorted(gemms.items(), key=lambda kv: -kv[1]["ms"])[:6]:
        a = v["flops"] / v["ms"] / 1e9
        inst.append({"tag": k, "launches_per_step": v["launches"] // nprof, "avg_launch_ms": v["ms"] / v["launches"],
                     "achieved": a, "frac": a / peak_tf, "share_of_library_kernel_time": v["ms"] / total_ms,
                     "traffic_ncu_bytes_per_launch": NCU_TRAFFIC_BYTES.get(k)})
    known = [(NCU_TRAFFIC_BYTES[k], v["launches"]) for k, v in gemms.items() if k in NCU_TRAFFIC_BYTES]
    traffic = (sum(t * n for t, n in known) / sum(n for _, n in known)) if known else None
    roofline = {"bound": "tensor", "kernel": f"gemm_tc_kernel<BN,A_MN,B_MN,EPI> ({len(gemms)} instantiations, {all_n // nprof} launches/step)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                "traffic_note": "DRAM read+write bytes per launch (ncu --set full, in situ at C2, profiles/r02_call2/ncu_in_situ_metrics.txt), "
                                "launch-weighted over the instantiations that were captured; per-instantiation values under instantiations[]",
                "peak_source": peak_src, "flops_per_launch": all_f / all_n, "avg_launch_ms": all_ms / all_n,
                "share_of_library_kernel_time": all_ms / total_ms, "instantiations": inst,
                "library_kernel_ms_per_step": total_ms / nprof, "step_ms_under_profiler_events": t_ms / nprof,
                "timing": timing or "CUDA events around every library kernel on its launch stream, eager single-stream pass after the timed region, "
                                    "GPU-side head start before each profiled step so launches are queued before the GPU needs them"}
    return roofline, kernels


def gpu_head_start(dev, ms):
    """Returns a callable that makes the GPU spin for ~`ms` milliseconds on the current stream (torch.cuda._sleep takes SM
    clock cycles); a no-op when ms <= 0 or the helper is missing.  Used only by the per-kernel profile pass."""
    import torch
    sleep = getattr(torch.cuda, "_sleep", None)
    if ms <= 0 or sleep is None:
        return lambda: None
    try:
        khz = float(torch.cuda.get_device_properties(dev).clock_rate)
    except Exception:
        khz = 0.0
    if not khz > 0:
        khz = 1.965e6                                    # B200 boost clock
    cycles = int(ms * khz)                               # kHz * ms = cycles
    return lambda: sleep(cycles)


def parse_profile(lib, graph_only=False):
    """graph_only: only the records created under stream capture (the library marks them with a leading '@')."""
    import ctypes as C
    buf = C.create_string_buffer(1 << 17)
    lib.fm_profile_report(buf, len(buf))
    rows = {}
    for line in buf.value.decode().splitlines():
        tag, n, ms, flops, byts = line.split()
        captured = tag.startswith("@")
        if graph_only and not captured:
            continue
        tag = tag.lstrip("@")
        o = rows.setdefault(tag, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        o["launches"] += int(n); o["ms"] += float(ms); o["flops"] += float(flops); o["bytes"] += float(byts)
    return rows


_KERNEL_TAG = {"ln_fwd_w_kernel": "ln_fwd", "ln_fwd_kernel": "ln_fwd", "ln_bwd_kernel": "ln_bwd", "ln_bwd_dx_w_kernel": "ln_bwd_dx",
               "ln_bwd_dgb_w_kernel": "ln_bwd_dgb", "ln_bwd_reduce_kernel": "ln_bwd_reduce", "xattn_core_fwd_tc_kernel": "xattn_core_fwd",
               "xattn_core_bwd_tc_kernel": "xattn_core_bwd", "resampler_core_fwd_tc_kernel": "resampler_core_fwd",
               "resampler_core_bwd_tc_kernel": "resampler_core_bwd", "cast_f32_bf16_kernel": "cast_f32_bf16", "ce_fwd_kernel": "ce_fwd",
               "ce_bwd_kernel": "ce_bwd", "alpha_grad_kernel": "alpha_grad", "group_rowsum_kernel": "group_rowsum",
               "bcast_rows_kernel": "bcast_rows", "text_time_kernel": "text_time", "dot_reduce_kernel": "dot_reduce", "adamw_kernel": "adamw"}


def kernel_family(name: str):
    """CUPTI (demangled) kernel name of a library kernel -> profiler tag family; None for foreign kernels."""
    import re
    if "fm::" not in name:
        return None
    m = re.search(r"gemm_tc_kernel<\(?(?:int\))?\s*(\d+),\s*\(?(?:bool\))?\s*(true|false|0|1),\s*\(?(?:bool\))?\s*(true|false|0|1),\s*\(?(?:int\))?\s*(\d+)>", name)
    if m:
        b = lambda t: 1 if t in ("true", "1") else 0      # noqa: E731
        return f"gemm_a{b(m.group(2))}b{b(m.group(3))}_epi{m.group(4)}_bn{m.group(1)}"
    base = re.search(r"fm::([A-Za-z0-9_]+)", name).group(1)
    return _KERNEL_TAG.get(base, base)


def tag_family(tag: str) -> str:
    t = tag.lstrip("@")
    t = t.split("/", 1)[1] if "/" in t else t
    return t[:-6] if t.endswith("_group") else t


def parse_launch_log(lib):
    """fm_profile_log -> [(scope, family, tag, flops, bytes)] in launch order (records made under stream capture only)."""
    import ctypes as C
    buf = C.create_string_buffer(1 << 18)
    if lib.fm_profile_log(buf, len(buf)) != 0:
        return []
    out = []
    for line in buf.value.decode().splitlines():
        tag, flops, byts = line.split()
        if not tag.startswith("@"):
            continue
        t = tag[1:]
        scope = t.split("/", 1)[0] if "/" in t else ""
        out.append((scope, tag_family(t), t.split("/", 1)[1] if "/" in t else t, float(flops), float(byts)))
    return out


def cupti_profile(replay, nprof, launch_log):
    """Device-side durations (CUPTI activity records through torch.profiler) of the library's kernels inside `nprof` replays of
    the timed CUDA graph, attributed to the library's own launch log: within one kernel family (= one template instantiation)
    every launch lives on one stream, so the k-th record of a family (by start time) is the k-th launch the library logged for it.
    Returns (prof dict tag -> {launches, ms, flops, bytes} with scope-prefixed tags, wall ms of the profiled replays) or None."""
    import collections
    from torch.profiler import ProfilerActivity, profile
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ev0.record()
        for _ in range(nprof):
            replay()
        ev1.record()
        torch.cuda.synchronize()
    per_fam = collections.defaultdict(list)
    for e in prof.events():
        fam = kernel_family(e.name)
        if fam is None:
            continue
        dur = getattr(e, "device_time", None) or getattr(e, "cuda_time", None) or 0.0
        if not dur:
            dur = getattr(e, "device_time_total", 0.0) or getattr(e, "cuda_time_total", 0.0)
        per_fam[fam].append((e.time_range.start, float(dur)))
    if not per_fam:
        return None
    logged = collections.defaultdict(list)
    for scope, fam, tag, flops, byts in launch_log:
        logged[fam].append((scope, tag, flops, byts))
    out = {}
    unmatched = []
    for fam, recs in per_fam.items():
        recs.sort()
        want = logged.get(fam, [])
        if want and len(recs) == nprof * len(want):
            for i, (_, dur) in enumerate(recs):
                scope, tag, flops, byts = want[i % len(want)]
                o = out.setdefault((scope + "/" if scope else "") + tag, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
                o["launches"] += 1; o["ms"] += dur / 1e3; o["flops"] += flops; o["bytes"] += byts
        else:      # cannot attribute launch by launch: family totals only (flops from the log if the counts allow it)
            unmatched.append((fam, len(recs), len(want)))
            o = out.setdefault(fam, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            o["launches"] += len(recs); o["ms"] += sum(d for _, d in recs) / 1e3
            o["flops"] += nprof * sum(w[2] for w in want); o["bytes"] += nprof * sum(w[3] for w in want)
    return out, ev0.elapsed_time(ev1), unmatched


def usable_cores() -> int:
    """Host threads this process may really use: affinity mask, capped by a cgroup CPU quota if one is set."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def time_cpu_reference(w, B, steps, warmup, threads=None, max_seconds=240.0):
    """fwd+bwd of the same step on the host cores: the oracle restatement of the two hot-path modules spliced into the stock HF
    LM (oracle/oracle_model.py; fp32; nothing of flamingo_mini_b200 is imported, so the product's .so is never mapped).
    Returns (samples/s, s/step, threads used, timed steps).  torch's intra-op pool scales poorly past ~32 threads on these
    matrix sizes, so a short probe on a batch-4 slice picks the fastest of {16, 32, 64, all usable cores}; the timed steps then
    run the FULL per-GPU batch.  `max_seconds` bounds the whole call: the number of timed steps shrinks (never below 1) if the
    first full step predicts an overrun."""
    from oracle.oracle_model import OracleFlamingo
    model = OracleFlamingo(w["lm"], w["lm_config"], w["D"], w["Dv"], xattn_every=w["xattn_every"]).float()
    clip, ids, ml = make_batch(w, B, "cpu", 1234, torch.float32)
    t_start = time.perf_counter()

    def one(nb=B):
        model.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        model.training_step(clip[:nb * w["N"]], ids[:nb], ml[:nb], w["N"])
        return time.perf_counter() - t0

    if threads is None:
        # ascending probe; stop as soon as more threads stop helping (oversubscribed pools can be >30x slower)
        cores = usable_cores()
        best = None
        nb = min(4, B)
        for cand in sorted({min(cores, 16), min(cores, 32), min(cores, 64), cores}):
            torch.set_num_threads(cand)
            if best is None:
                one(nb)                    # first touch / lazy init
            t = one(nb)
            if best is None or t < best[0]:
                best = (t, cand)
            if t > 1.1 * best[0] or t > 20.0:
                break
        threads = best[1]
    torch.set_num_threads(threads)
    times = []
    first = one()                          # always one untimed full-batch step (allocator / thread-pool warm-up)
    budget = max_seconds - (time.perf_counter() - t_start)
    warm_left = max(0, warmup - 1)
    if (warm_left + steps) * first > budget:
        warm_left = 0
        steps = max(1, min(steps, int(budget / first)))
    for i in range(warm_left + steps):
        t = one()
        if i >= warm_left:
            times.append(t)
    t = sum(times) / len(times)
    return B / t, t, threads, len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="batch of the CPU-baseline sample (0 = the workload's full per-GPU batch)")
    ap.add_argument("--reference-max-seconds", type=float, default=240.0, help="--impl reference: bound on the whole run (timed steps shrink to fit)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--profile-multi", action="store_true", help="also run the per-kernel profile pass when N > 1")
    ap.add_argument("--profile-head-start-ms", type=float, default=40.0,
                    help="GPU-side spin ahead of each profiled (eager, per-kernel events) step so launches are queued before the GPU needs them; 0 = off")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--torch-loss", action="store_true", help="loss head through torch's cross_entropy instead of the library's row kernels")
    ap.add_argument("--fused-loss", action="store_true",
                    help="loss head through fm_cross_entropy_{fwd,bwd} instead of torch's (default; --torch-loss selects torch's)")
    ap.add_argument("--per-layer-reduce", action="store_true", help="(default since round 2; kept for old command lines)")
    ap.add_argument("--split-embedding", action="store_true", help="(default since round 2; kept for old command lines)")
    ap.add_argument("--reduce-bucket", type=int, default=1,
                    help="N>1: gradient arenas of this many consecutive gated xattn blocks share one buffer and one all-reduce")
    ap.add_argument("--bf16-wire", action="store_true",
                    help="N>1: fp32 gradient arenas cross NVLink as bf16 (GradArenaReducer.wire_dtype; like DDP's bf16_compress_hook). "
                         "Off by default: on 2 B200s the two cast passes cost more than the halved collective saves "
                         "(10.99 vs 10.62 ms/step, profiles/r02_call9_dp2)")
    ap.add_argument("--whole-arena-reduce", action="store_true",
                    help="N>1: all-reduce the resampler's gradient arena in one piece after its backward instead of layer by layer "
                         "during it (fm_resampler_bwd_notify)")
    ap.add_argument("--dense-embedding-reduce", action="store_true",
                    help="N>1: one dense all-reduce of the tied token-embedding gradient after backward instead of an early dense "
                         "all-reduce of the lm_head part + gathered lookup rows (parallel.SplitEmbeddingGrad)")
    args = ap.parse_args()
    # measured on 2 B200s (profiles/r02_call4_dp): default exchange 11.17 ms/step, split embedding 10.92, per-layer 11.03, both 10.76
    args.per_layer_reduce = not args.whole_arena_reduce
    args.split_embedding = not args.dense_embedding_reduce
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cores = os.cpu_count() or 1
    metric = "image-text samples/sec (fwd+bwd)"
    config = {"workload": f"{args.workload}: {w['desc']}", "global_batch": w["B"] * world, "seq_len": w["S"],
              "parallelism": f"dp{world}", "clip_tokens": w["F"], "images_per_sample": w["N"],
              "optimizer_step": "none (metric is fwd+bwd)",
              "frozen_lm": "stock HF GPT-2/OPT, train mode (dropout on); gelu_new evaluated by torch's single tanh-GELU "
                           "kernel (FlamingoConfig.lm_fused_gelu, same formula, both arms)",
              "l2": "per-step working set (weights + activations > 1 GB) exceeds the 126 MB L2; no explicit flush"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        Bs = w["B"] if args.cpu_sample_batch <= 0 else min(args.cpu_sample_batch, w["B"])
        sps, t, cores, timed = time_cpu_reference(w, Bs, max(args.steps, 1), args.warmup, max_seconds=args.reference_max_seconds)
        line = {"impl": "reference", "metric": metric, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": timed, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": f"oracle port (oracle/oracle_model.py) + stock HF LM, fp32, batch {Bs} of {w['B']} per step, same seq/config, "
                                           f"{timed} timed steps of {args.steps} requested (bounded to {args.reference_max_seconds:.0f} s), "
                                           f"torch threads={cores} of {usable_cores()} usable"},
                "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    from flamingo_mini_b200 import _lib
    from flamingo_mini_b200.parallel import GradArenaReducer, hot_path_modules
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the sm_100a library has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    config["library"] = os.path.basename(_lib.lib_path())
    if os.environ.get("FM_B200_OPTS"):
        config["library_options"] = os.environ["FM_B200_OPTS"]
    model = build_model(w, dev, fused_loss=not args.torch_loss)
    if not args.torch_loss:
        config["loss_head"] = "fm_cross_entropy_fwd/bwd (library row kernels)"
    hot = hot_path_modules(model)
    hot_ids = {id(p) for m in hot for p in m.parameters()}
    extra = [p for p in model.parameters() if p.requires_grad and id(p) not in hot_ids]
    reducer = (GradArenaReducer(hot, extra_params=extra, per_layer=args.per_layer_reduce,
                                wire_dtype=torch.bfloat16 if args.bf16_wire else None, bucket_blocks=args.reduce_bucket)
               if world > 1 else None)
    if reducer is not None and args.reduce_bucket > 1:
        config["grad_reduce_bucket_blocks"] = args.reduce_bucket
    if reducer is not None:
        config["grad_wire_dtype"] = ("bf16 (fp32 arenas are rounded to bf16 for the all-reduce and restored into the fp32 arena afterwards)"
                                     if args.bf16_wire else "fp32")
    if reducer is not None and args.per_layer_reduce:
        config["resampler_grad_exchange"] = "per layer" if _lib.has("fm_resampler_bwd_notify") else "whole arena (entry point not in this build)"
    if reducer is not None and args.split_embedding:
        from flamingo_mini_b200.parallel import SplitEmbeddingGrad
        SplitEmbeddingGrad.install(model, reducer)
        config["embedding_grad_exchange"] = "split (dense lm_head part all-reduced early, lookup rows all-gathered)"
    B = w["B"]
    clip, ids, ml = make_batch(w, B, dev, 1234 + rank, torch.bfloat16)
    for m in hot:                    # a real training step re-casts the fp32 masters to bf16 after every update
        m._fp.always_refresh = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph = {"g": None, "loss": None, "log": []}

    def step_eager():
        model.zero_grad(set_to_none=True)
        return train_step(model, w, clip, ids, ml, reducer)

    def capture_graph(warm=3, log=False):
        """Whole step (fwd + bwd + gradient all-reduce) as one CUDA graph: removes the host launch overhead of the
        ~1k kernels per step (stock HF LM included).  Inputs live in the static tensors clip/ids/ml."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warm):
                step_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        model.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        if log and _lib.has("fm_profile_log"):
            lib.fm_profile_enable(2)          # launch log only: which tag / FLOPs / module scope each captured launch has
        with torch.cuda.graph(g):
            loss = train_step(model, w, clip, ids, ml, reducer)
        if log and _lib.has("fm_profile_log"):
            graph["log"] = parse_launch_log(lib)
            lib.fm_profile_enable(0)
        return g, loss

    def capture():
        graph["g"], graph["loss"] = capture_graph(log=True)

    def run(n, e2e=False, host=None):
        loss = None
        for _ in range(n):
            if e2e:      # H2D of this step's inputs from pinned memory, D2H of its loss
                clip.copy_(host[0], non_blocking=True); ids.copy_(host[1], non_blocking=True); ml.copy_(host[2], non_blocking=True)
            if graph["g"] is not None:
                graph["g"].replay()
                loss = graph["loss"]
            else:
                loss = step_eager()
            if e2e:
                host[3].copy_(loss.detach().float(), non_blocking=False)
        return loss

    def timed(n, **kw):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        loss = run(n, **kw)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), loss

    used_graph = False
    if not args.no_graph:
        try:
            capture()
            used_graph = True
        except Exception as e:           # fall back to eager launches, and say so in the JSON line
            graph["g"] = None
            config["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:300]
            torch.cuda.synchronize()
    config["cuda_graph"] = used_graph
    launches_per_step = None
    if used_graph:                        # kernels inside a replayed graph are not counted by the library's host counter
        l0 = lib.fm_launch_count(); step_eager(); launches_per_step = lib.fm_launch_count() - l0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()               # nvidia-smi needs ~1 s before its first sample: start ahead of the warm-up
    run(warmup)
    launches0 = lib.fm_launch_count()
    t_begin = time.time()
    ms, loss = timed(args.steps)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    launches = lib.fm_launch_count() - launches0
    if used_graph:
        launches = launches_per_step * args.steps
    value = B * world * args.steps / (ms / 1e3)

    # end to end: pinned host inputs -> H2D each step, loss D2H each step
    host = [clip.cpu().pin_memory(), ids.cpu().pin_memory(), ml.cpu().pin_memory(), torch.zeros((), dtype=torch.float32).pin_memory()]
    if graph["g"] is not None:
        graph["loss"] = graph["loss"].detach()
    run(2, e2e=True, host=host)
    ms_e2e, _ = timed(args.steps, e2e=True, host=host)
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host[:3])

    # per-kernel timing, in order of preference: (1) CUPTI records of a second, single-stream capture of the step attributed
    # through the library's launch log; (2) fallback: external event-record nodes around every library kernel inside such a
    # capture (over-reads each kernel by a few microseconds of graph-node latency); (3) without a graph: eager pass with events.
    roofline, kernels, xattn = None, None, None
    if world > 1 and not args.profile_multi:
        config["profile"] = "per-kernel profile is taken at N=1 (identical kernels per rank under data parallelism); --profile-multi forces it"
    if not args.no_profile and (world == 1 or args.profile_multi):
        lib.fm_set_option(0, 0)            # per-kernel event timing needs one stream: no side-stream overlap in this pass
        lib.fm_profile_enable(1)
        nprof = min(3, args.steps)
        t_ms = 0.0
        prof = {}

        def accumulate(rows):
            for k, v in rows.items():
                o = prof.setdefault(k, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
                for f in o:
                    o[f] += v[f]

        timing = None
        prof_graph = None
        cupti = None
        if used_graph and _lib.has("fm_profile_log"):
            # A SECOND capture of the same step with the library's side stream off (fm_set_option above) and its launch log on:
            # with one stream no two library kernels overlap, so a CUPTI duration is the kernel's own (two persistent
            # one-CTA-per-SM GEMMs that overlap stretch each other: in the timed graph dy1n = dh W1 reads 45 us next to the
            # side-stream dW GEMMs and 25 us alone) and the sum over kernels cannot exceed the replay's duration.
            lib.fm_profile_enable(0)
            try:
                prof_graph, _ = capture_graph(warm=1, log=True)
                barrier()
                cupti = cupti_profile(prof_graph.replay, nprof, graph["log"])
                barrier()
            except Exception as e:
                config["cupti_profile_error"] = f"{type(e).__name__}: {e}"[:200]
                cupti = None
                torch.cuda.synchronize()
            prof_graph = None
            lib.fm_profile_enable(1)
        if cupti is not None:
            prof, t_ms, unmatched = cupti
            if unmatched:
                config["cupti_unattributed_families"] = unmatched[:8]
            timing = ("device-side kernel durations (CUPTI activity records via torch.profiler) of the library's kernels inside replays of a "
                      "CUDA graph of the same step captured with the library's side stream off (no two library kernels overlap), attributed "
                      "launch by launch through the library's own launch log (fm_profile_log); no host latency, no event-node overhead "
                      "inside the intervals")
        elif used_graph:
            try:
                prof_graph, _ = capture_graph(warm=1)
            except Exception as e:
                config["profile_graph_error"] = f"{type(e).__name__}: {e}"[:200]
                prof_graph = None
                torch.cuda.synchronize()
                lib.fm_profile_enable(1)       # drop the eager warm-up records
        if cupti is not None:
            pass
        elif prof_graph is not None:
            for _ in range(nprof):
                barrier()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                prof_graph.replay()
                ev1.record()
                barrier()
                t_ms += ev0.elapsed_time(ev1)
                accumulate(parse_profile(lib, graph_only=True))
            timing = ("CUDA events recorded as external event-record nodes around every library kernel inside a CUDA-graph capture of the "
                      "step (single stream, side stream off); durations are those of the replayed graph")
            del prof_graph
            torch.cuda.synchronize()
        else:
            saved_graph, graph["g"] = graph["g"], None        # eager pass with a GPU-side head start (see gpu_head_start)
            lib.fm_profile_enable(1)
            head_start = gpu_head_start(dev, args.profile_head_start_ms)
            for _ in range(nprof):
                barrier()
                head_start()                                  # enqueued, not waited for: the host runs ahead from here
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                run(1)
                ev1.record()
                barrier()
                t_ms += ev0.elapsed_time(ev1)
            graph["g"] = saved_graph
            accumulate(parse_profile(lib))
        lib.fm_profile_enable(0)
        lib.fm_set_option(0, 1)
        roofline, kernels = make_roofline(prof, nprof, t_ms, timing=timing)
        if roofline is not None:
            xs = scope_totals(prof, nprof, "x")
            n_blocks = len(range(0, w["lm_config"].get("n_layer", w["lm_config"].get("num_hidden_layers")), w["xattn_every"]))
            I, Q = 512, 64
            fwd_x = 2 * w["S"] * w["D"] * I + 4 * (w["N"] * Q) * w["Dv"] * I + 4 * w["S"] * 64 * I + 2 * w["S"] * I * w["D"] + 16 * w["S"] * w["D"] ** 2
            alg = 3.0 * fwd_x * n_blocks * B                 # bwd_X = 2 fwd_X, masked-useful attention FLOPs only
            if xs["ms_per_step"] > 0:
                tf = alg / xs["ms_per_step"] / 1e9
                xattn = {"tflops": tf, "peak": roofline["peak"], "frac": tf / roofline["peak"], "unit": "TFLOP/s",
                         "algorithmic_gflop_per_step": alg / 1e9, "kernel_ms_per_step": xs["ms_per_step"],
                         "launches_per_step": xs["launches_per_step"],
                         "definition": "sum over gated xattn blocks of (fwd_X + bwd_X) * batch / time in ALL kernels launched by fm_xattn_fwd/bwd "
                                       "(GEMMs, attention cores, LayerNorm, casts)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bs = w["B"] if args.cpu_sample_batch <= 0 else min(args.cpu_sample_batch, w["B"])
        sps, t, used, timed = time_cpu_reference(w, Bs, 2, 0, max_seconds=45.0)
        cpu_baseline = {"value": sps, "unit": "samples/s", "cores": used, "kind": "port",
                        "sample": f"oracle port + stock HF LM, fp32, batch {Bs} of {w['B']}, thread-count probe + {timed} timed steps ({t:.2f} s/step), {usable_cores()} usable cores"}

    if rank == 0:
        fl = hot_path_flops_per_sample(w)
        line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "xattn": xattn, "cpu_baseline": cpu_baseline,
                "hot_path": {"gflop_per_sample_fwd_bwd": fl / 1e9,
                             "library_kernel_ms_per_step": (roofline or {}).get("library_kernel_ms_per_step"),
                             "tflops_over_library_kernel_time": (fl * B / ((roofline or {}).get("library_kernel_ms_per_step") or float("nan")) / 1e9)},
                "kernels": kernels, "loss": float(loss.detach())}
        print(json.dumps(line), flush=True)
    # Tear down in dependency order: captured graphs (they hold NCCL work) first, then the process group.  A watchdog turns a
    # teardown that does not return (seen once in round 1 with a live graph holding collectives) into a clean exit instead of
    # a hung multi-GPU box: the measurements are already printed and synchronised at this point.
    torch.cuda.synchronize()
    if world > 1:
        import gc
        dist.barrier()
        sys.stdout.flush()
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        graph["g"] = None
        graph["loss"] = None
        del model, reducer
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        watchdog.cancel()
    return 0


if __name__ == "__main__":
    sys.exit(main())
