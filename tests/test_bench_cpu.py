"""CPU checks of bench.py's host logic: FLOP table vs BASELINE.md, roofline assembly, reference-arm JSON contract."""
import json
import subprocess
import sys

import bench


def test_flop_table_matches_baseline_md():
    # BASELINE.md "Algorithmic FLOPs per sample": C2 70.7, C3 846.8, C5 6 841 GFLOP
    assert abs(bench.hot_path_flops_per_sample(bench.WORKLOADS["c2"]) / 1e9 - 70.7) < 0.1
    assert abs(bench.hot_path_flops_per_sample(bench.WORKLOADS["c3"]) / 1e9 - 846.8) < 0.1
    assert abs(bench.hot_path_flops_per_sample(bench.WORKLOADS["c5"]) / 1e9 - 6841) < 1


def test_make_roofline():
    prof = {"x/gemm_a0b0_epi1_bn256": dict(launches=36, ms=1.2, flops=36 * 19.3e9, bytes=1e9),
            "x/gemm_a1b1_epi0_bn64": dict(launches=100, ms=1.3, flops=100 * 3.2e9, bytes=0.6e9),
            "r/gemm_a1b1_epi0_bn64": dict(launches=62, ms=0.8, flops=62 * 3.2e9, bytes=0.4e9),
            "ln_bwd": dict(launches=129, ms=1.8, flops=0.0, bytes=4e9)}
    r, k = bench.make_roofline(prof, 3, 60.0, peaks_path="/nonexistent")
    assert r["bound"] == "tensor" and r["peak"] == 1400.0 and "fallback" in r["peak_source"]
    want = (36 * 19.3e9 + 162 * 3.2e9) / 3.3 / 1e9
    assert abs(r["achieved"] - want) < 1e-6 and abs(r["frac"] - want / 1400.0) < 1e-9
    assert r["instantiations"][0]["tag"] == "gemm_a1b1_epi0_bn64"
    assert r["instantiations"][0]["traffic_ncu_bytes_per_launch"] is None            # not captured with ncu
    assert r["instantiations"][1]["traffic_ncu_bytes_per_launch"] == bench.NCU_TRAFFIC_BYTES["gemm_a0b0_epi1_bn256"]
    assert r["traffic"] == bench.NCU_TRAFFIC_BYTES["gemm_a0b0_epi1_bn256"]
    assert bench.scope_totals(prof, 1, "x")["launches_per_step"] == 136
    assert k["ln_bwd"]["launches"] == 43 and k["ln_bwd"]["tflops"] is None
    assert bench.make_roofline({"ln_fwd": dict(launches=3, ms=0.1, flops=0.0, bytes=1.0)}, 3, 1.0)[0] is None


def test_gpu_head_start_is_inert_without_a_request():
    """The profile pass's GPU-side head start: off at 0 ms, and building the callable never touches a device."""
    calls = []
    assert bench.gpu_head_start("cpu", 0.0)() is None
    import torch
    orig = torch.cuda._sleep
    torch.cuda._sleep = lambda cycles: calls.append(cycles)
    try:
        bench.gpu_head_start("cpu", 2.0)()               # no CUDA device here: falls back to the B200 boost clock
    finally:
        torch.cuda._sleep = orig
    assert calls == [int(2.0 * 1.965e6)]


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=bench.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_never_maps_the_product_library():
    """`--impl reference` times the oracle port + stock HF only: neither flamingo_mini_b200 nor its .so may be loaded."""
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'tiny', '--steps', '1', '--warmup', '0']\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libflamingo_b200' not in maps, 'product .so mapped by the reference arm'\n"
            "assert not any(m.startswith('flamingo_mini_b200') for m in sys.modules), 'product package imported'\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=bench.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]


def test_cupti_name_to_tag_mapping():
    """bench.py attributes CUPTI kernel records to the library's profiler tags by template instantiation."""
    f = bench.kernel_family
    assert f("void fm::gemm_tc_kernel<256, false, false, 1>(fm::GemmGroup)") == "gemm_a0b0_epi1_bn256"
    assert f("void fm::gemm_tc_kernel<(int)128, (bool)1, (bool)1, (int)0>(fm::GemmGroup)") == "gemm_a1b1_epi0_bn128"
    assert f("void fm::ln_fwd_w_kernel<3>(fm::LnArgs)") == "ln_fwd" and f("void fm::ln_bwd_kernel<64, 2>(fm::LnBwdArgs)") == "ln_bwd"
    assert f("fm::xattn_core_bwd_tc_kernel(CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, fm::XTcBwdArgs)") == "xattn_core_bwd"
    assert f("void at::native::vectorized_elementwise_kernel<8, ...>") is None
    assert bench.tag_family("@x/gemm_a1b1_epi0_bn64_group") == "gemm_a1b1_epi0_bn64" and bench.tag_family("r/ln_fwd") == "ln_fwd"
