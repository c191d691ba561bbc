"""GPU parity of the PREPARED, opt-in cooperative-SH forward preprocess (PGS_SH_COOP=1, csrc/preprocess_fwd.cu:
preprocess_fwd_coop_kernel) against the unmodified reference CUDA build — the same bit-exact stage checks as the default
kernel's, rerun with the switch on, for the point-level and the block-level (fused superquadric) instantiation.

STATUS: the variant was written with this round's GPU budget spent.  On the CPU emulator its forward outputs are
bit-identical to the default kernel's and its SASS has the default's FP opcode mix (tests/test_emu_raster.py,
tests/test_abi.py), but it has not yet run on a B200 -> non-strict xfail (a pass shows as XPASS), sorted after every
other GPU suite.  The default path does not depend on it.  Round 2: read the result, time it (tools/gpu_round2_first.sh),
then either make it the default or delete it."""
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300),
              pytest.mark.xfail(strict=False, reason="PGS_SH_COOP=1: bit-identical on the CPU emulator, first run on a "
                                                     "B200 pending")]


@pytest.mark.parametrize("name,P", [("C1", None), ("C3", 200_000)])
def test_cooperative_sh_forward_stages_bit_exact_vs_reference(monkeypatch, name, P):
    import test_gpu_base_raster as t
    from partgs_b200 import _lib
    monkeypatch.setenv("PGS_SH_COOP", "1")
    n0 = _lib.load().pgs_launch_count()
    t.test_forward_stages_bit_exact_vs_reference(name, P)
    assert _lib.load().pgs_launch_count() > n0


def test_cooperative_sh_backward_and_lower_degrees_vs_reference(monkeypatch):
    import test_gpu_base_raster as t
    monkeypatch.setenv("PGS_SH_COOP", "1")
    t.test_backward_vs_reference("C2", None)
    t.test_lower_sh_degree_vs_reference()


def test_cooperative_sh_in_the_fused_block_kernel(monkeypatch):
    import test_gpu_blocks_fused as t
    monkeypatch.setenv("PGS_SH_COOP", "1")
    t.test_fused_blocks_match_unfused_composition(5, 3, 333, 201)   # both sides then run a cooperative instantiation
