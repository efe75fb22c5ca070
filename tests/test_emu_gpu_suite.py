"""Runs the bodies of the `-m gpu` tests on the CPU against tests/simt_emu (SVAE_TEST_BACKEND=emu, see
tests/conftest.py): the very tests the B200 box executes at round end, minus the ones that are too large for a
fiber-based emulation or that need real CUDA machinery (CUDA graphs, full-size shapes).
One subprocess, so the emulation patches never leak into the other CPU tests."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# too large for the emulation (minutes: the fp32 SIMT GEMM at H = 500), or CUDA graphs / events of torch.cuda (they do not exist on a host).
# The tcgen05 GEMM unit tests and the C1-shape step in FAST precision (H = 500, two n-tiles, fused output dot) DO run:
# tc_gemm.cu itself executes on the host model of tcgen05 / TMA / mbarriers (tests/simt_emu/tc_emu.h).
DESELECT = ("not (c1_shape and parity) and not c2_full_size and not full_model_size and not graphed_step "
            "and not baseline_model_shapes and not in_kernel_eps and not trajectory_at_c1 and not decoder_grads_event and not small_and_large_batch")


def test_gpu_test_bodies_pass_on_the_simt_emulation():
    from tests.simt_emu.build import build
    build()                                 # once, here: the workers below must not compile it concurrently
    env = dict(os.environ, SVAE_TEST_BACKEND="emu")
    env.pop("SVAE_CTF_FAST", None)
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_api.py", "tests/test_gpu_parity.py",
           "tests/test_gpu_zy_late.py", "tests/test_gpu_zz_options.py", "tests/test_gpu_round2.py", "-m", "gpu", "-q", "-p", "no:cacheprovider", "-k", DESELECT]
    try:                                    # four workers when pytest-xdist is available (each loads its own library)
        import xdist  # noqa: F401
        cmd += ["-n", "4"]
    except ImportError:
        cmd += ["-x"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in tail and " failed" not in tail and " skipped" not in tail, tail


import pytest  # noqa: E402


@pytest.mark.parametrize("sms,cta_group", [("4", "2"), ("3", "1")])
def test_tcgen05_gemm_persistent_paths_on_a_small_emulated_device(sms, cta_group):
    """The tcgen05 GEMM unit tests (tc_gemm.cu and the fused backward kernels of tc_bwd.cu) again on an emulated device with only a few SMs, so that every persistent CTA
    (pair) walks several tiles: accumulator double buffering, ring wrap-around and mbarrier phase flips are exercised
    (a wrong arrival count or a missing phase flip deadlocks here and the scheduler aborts with the waiters listed).
    cta_group 2 = CTA pairs (the default), 1 = the single-CTA variant (SVAE_TC_CTA_GROUP=1)."""
    from tests.simt_emu.build import build
    build()
    env = dict(os.environ, SVAE_TEST_BACKEND="emu", SVAE_EMU_SMS=sms, SVAE_TC_CTA_GROUP=cta_group)
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
           "-k", "tc_ and not 78400"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail
