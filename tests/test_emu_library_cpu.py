"""The STAGING library (csrc_next/, never run on a B200 yet) executed on the host emulator: the bodies of the -m gpu tests
are reused unchanged with DEV = "cpu" and the emulated libflamingo_b200_emu.so swapped in (tests/_emu_util.py).
What a pass means: kernel logic — tile schedules, mbarrier protocols, TMA boxes, UMMA descriptors, TMEM addressing,
epilogue indexing, reductions — is right under the functional model of tests/cpu_harness/tc_emu.h.  What it does not
mean: anything about speed, or about hardware behaviour outside that model."""
import pytest
import torch

from tests import _emu_util

pytestmark = pytest.mark.skipif(not _emu_util.available(), reason="needs g++ and the CUDA headers")


@pytest.fixture()
def emu(monkeypatch):
    import tests.test_gpu_gemm as G
    import tests.test_gpu_modules as M
    import tests.test_gpu_ops as P
    for mod in (G, M, P):
        monkeypatch.setattr(mod, "DEV", "cpu")
    with _emu_util.swapped_in() as lib:
        yield lib


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 0), (256, 256, 256, 64), (256, 256, 256, 192), (1000, 520, 200, 0),
                                      (1000, 520, 200, 128), (1000, 520, 200, 256), (136, 8, 72, 0)])
def test_gemm_store_f32(emu, a_mn, b_mn, M, N, K, bn):
    import tests.test_gpu_gemm as G
    G.test_gemm_store_f32(a_mn, b_mn, M, N, K, bn)


def test_gemm_epilogues(emu):
    import tests.test_gpu_gemm as G
    G.test_gemm_store_bf16_scale_bias_gate()
    for act in (0, 1, 2):
        G.test_gemm_act_epilogue(act)
    for aux_f32, out_f32 in [(0, 1), (1, 1), (1, 0), (0, 0)]:
        G.test_gemm_resid_epilogue(aux_f32, out_f32)
    G.test_gemm_dact_epilogue()
