"""csrc/extract.cu compiled for the CPU lock-step emulator (tests/cuda_emu) against the golden vectors of the
reference's own partmap_to_rgbmap, and against numpy for the unit normals."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "cuda_emu"))
import build as emu_build  # noqa: E402

Z = np.load(Path(__file__).parent / "golden" / "extract_maps.npz")
EXPORTS = r'''
extern "C" void emu_extract(int npix, int S, const float* semantic, const float* palette, int stride,
                            const float* rend_normal, float* part_rgb, float* normal_unit) {
  pgs::launch_extract_maps(npix, S, semantic, palette, stride, rend_normal, part_rgb, normal_unit, nullptr);
}
'''


@pytest.fixture(scope="module")
def emu():
    try:
        return C.CDLL(str(emu_build.build("extract.cu", EXPORTS)))
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_emulated_epilogue_matches_reference_golden(emu, name):
    part = np.ascontiguousarray(Z[f"{name}_part"]); pal = np.ascontiguousarray(Z[f"{name}_palette"])
    S, H, W = part.shape
    g = np.random.default_rng(1)
    nrm = g.normal(size=(3, H, W)).astype(np.float32)
    nrm[:, 0, 0] = 0
    rgb = np.full((3, H, W), np.nan, np.float32); unit = np.full((3, H, W), np.nan, np.float32)
    emu.emu_extract(H * W, S, _p(part), _p(pal), 3, _p(nrm), _p(rgb), _p(unit))
    assert np.array_equal(rgb, Z[f"{name}_rgb"])
    want = nrm / np.maximum(np.sqrt((nrm.astype(np.float64) ** 2).sum(0)), 1e-12)
    np.testing.assert_allclose(unit, want, rtol=2e-6, atol=1e-7)
    assert not unit[:, 0, 0].any()
    # halves can be skipped independently; an RGBA palette (stride 4) reads the same colours
    rgb2 = np.full((3, H, W), np.nan, np.float32)
    pal4 = np.concatenate([pal, np.ones((S + 1, 1), np.float32)], axis=1)
    emu.emu_extract(H * W, S, _p(part), _p(np.ascontiguousarray(pal4)), 4, None, _p(rgb2), None)
    assert np.array_equal(rgb2, rgb)


def test_emulated_epilogue_nan_semantics(emu):
    # torch.argmax treats NaN as the maximum (first one wins); a NaN sum is not "< 0.1"
    part = np.array([[[0.2, np.nan, 0.01]], [[0.9, 0.3, 0.02]], [[0.9, np.nan, 0.03]]], np.float32)  # [3,1,3]
    pal = np.arange(12, dtype=np.float32).reshape(4, 3)
    rgb = np.zeros((3, 1, 3), np.float32)
    emu.emu_extract(3, 3, _p(part), _p(pal), 3, None, _p(rgb), None)
    assert rgb[:, 0, 0].tolist() == pal[1].tolist()      # tie 0.9 / 0.9 -> first
    assert rgb[:, 0, 1].tolist() == pal[0].tolist()      # first NaN
    assert rgb[:, 0, 2].tolist() == [1.0, 1.0, 1.0]      # sum 0.06 < 0.1 -> white
