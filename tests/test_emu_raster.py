"""The WHOLE product library (every csrc/*.cu incl. the C-ABI layer) compiled for the CPU lock-step emulator
(tests/cuda_emu) and driven through its real C ABI (pgs_dsr_forward / pgs_dsr_backward) with host buffers, against
the C restatement of the reference (oracle/surfel_oracle.c).  A GPU-less functional regression test of preprocess,
scan, key duplication, the onesweep sort, tile ranges, both render kernels and the preprocess backward; the bit-exact
hardware parity against the reference CUDA build is in tests/test_gpu_*.py.  (glibc transcendentals and no FMA
contraction here, so values agree to rounding, not bit for bit.)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).parent / "cuda_emu"))
import build as emu_build  # noqa: E402

from oracle import cpu_oracle  # noqa: E402
from partgs_b200 import _lib, synth  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    try:
        lib = C.CDLL(str(emu_build.build_full()))
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)          # every ABI symbol exists in the emulator build too
        fn.restype, fn.argtypes = res, args
    return lib


class HostAlloc:
    """The three resize callbacks of the ABI, backed by numpy buffers."""
    def __init__(self):
        self.bufs = {}
        self.cb = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, nbytes, user):
        buf = np.zeros(int(nbytes) + 256, np.uint8)
        self.bufs[int(user or 0)] = buf
        return (buf.ctypes.data + 255) // 256 * 256

    def ptr(self, slot):
        return (self.bufs[slot].ctypes.data + 255) // 256 * 256

    def nbytes(self, slot):
        return self.bufs[slot].size - 256


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_emulated(emu, scene, cam, g_color, g_allmap, degree=3, colors_precomp=None, bg=None, scale_modifier=1.0):
    f32 = lambda t: np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)
    m3, sc, rot, op, sh = (f32(scene[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs"))
    vm, pm, cp = f32(cam.viewmatrix), f32(cam.projmatrix), f32(cam.campos)
    bg = np.zeros(3, np.float32) if bg is None else np.ascontiguousarray(bg, dtype=np.float32)
    P, M = m3.shape[0], sh.shape[1]
    cpre = None if colors_precomp is None else f32(colors_precomp)
    sh_arg = None if cpre is not None else sh
    W, H = cam.image_width, cam.image_height
    color = np.full((3, H, W), np.nan, np.float32)
    allmap = np.full((7, H, W), np.nan, np.float32)
    radii = np.full(P, -7, np.int32)
    al = HostAlloc()
    R = emu.pgs_dsr_forward(al.cb, 1, al.cb, 2, al.cb, 3, P, degree, M, _p(bg), W, H, _p(m3), _p(sh_arg), _p(cpre), _p(op), _p(sc),
                            scale_modifier, _p(rot), None, _p(vm), _p(pm), _p(cp), cam.tanfovx, cam.tanfovy, 0, _p(color),
                            _p(allmap), _p(radii), 1, None)
    assert R >= 0, emu.pgs_last_error()
    lay = _lib.DsrLayout()
    assert emu.pgs_dsr_get_layout(P, W, H, al.nbytes(2), C.byref(lay)) == 0, emu.pgs_last_error()
    base = al.ptr(2)
    point_list = np.ctypeslib.as_array(C.cast(base + lay.binning_point_list, C.POINTER(C.c_uint32)), (max(R, 1),))[:R].copy()
    keys = np.zeros(max(R, 1), np.uint64)     # rebuilt on request: the production path never materialises them
    assert emu.pgs_dsr_sorted_keys(P, W, H, al.ptr(1), base, al.nbytes(2), R, _p(keys), None) == 0, emu.pgs_last_error()
    keys = keys[:R].copy()
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    ranges = np.ctypeslib.as_array(C.cast(al.ptr(3) + lay.image_ranges, C.POINTER(C.c_uint32)), (ntiles, 2)).copy()
    g = dict(means2D=np.full((P, 3), np.nan, np.float32), colors=np.full((P, 3), np.nan, np.float32),
             opacity=np.full((P, 1), np.nan, np.float32), means3D=np.full((P, 3), np.nan, np.float32),
             transMat=np.full((P, 9), np.nan, np.float32), sh=np.full((P, M, 3), np.nan, np.float32),
             scales=np.full((P, 2), np.nan, np.float32), rotations=np.full((P, 4), np.nan, np.float32))
    scratch = np.zeros(emu.pgs_dsr_backward_scratch_bytes(P) + 256, np.uint8)
    gc, ga = f32(g_color), f32(g_allmap)
    rc = emu.pgs_dsr_backward(P, degree, M, R, _p(bg), W, H, _p(m3), _p(sh_arg), _p(cpre), _p(sc), scale_modifier, _p(rot), None, _p(vm),
                              _p(pm), _p(cp), cam.tanfovx, cam.tanfovy, _p(radii), al.ptr(1), al.ptr(2), al.nbytes(2),
                              al.ptr(3), _p(gc), _p(ga), _p(g["means2D"]), (scratch.ctypes.data + 255) // 256 * 256,
                              _p(g["opacity"]), _p(g["colors"]), _p(g["means3D"]), _p(g["transMat"]), _p(g["sh"]),
                              _p(g["scales"]), _p(g["rotations"]), 1, None)
    assert rc >= 0, emu.pgs_last_error()
    return dict(R=R, color=color, allmap=allmap, radii=radii, point_list=point_list, keys=keys, ranges=ranges, grads=g)


def rel(a, b, q=1.0):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    d = np.abs(a - b)
    return float((np.quantile(d, q) if q < 1 else d.max()) / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("P,W,H,seed", [(300, 48, 32, 1), (500, 40, 24, 2)])   # second: ragged right / bottom tiles
def test_emulated_forward_backward_match_the_c_oracle(emu, P, W, H, seed):
    scene = synth.make_point_scene(P, seed=seed, device="cpu")
    scene["scales"] = scene["scales"] * 3.0          # splats of a few pixels at this tiny resolution
    cam = synth.make_cameras(1, W, H, seed=seed + 10, device="cpu")[0]
    g = synth.upstream_grads(W, H, 5, device="cpu")
    ours = run_emulated(emu, scene, cam, g["color"], g["allmap"])
    f = cpu_oracle.forward_scene(scene, cam, keep_state=True)
    gr = cpu_oracle.backward(f, g["color"], g["allmap"])
    assert ours["R"] > 50 and (ours["radii"] > 0).sum() > 20, "scene does not exercise the rasteriser"
    # integer state: exact (both sides compute in IEEE fp32 without contraction here)
    assert np.array_equal(ours["radii"], f["radii"])
    assert ours["R"] == f["num_rendered"]
    assert np.array_equal(ours["keys"].astype(np.int64), f["keys"])
    assert np.array_equal(ours["point_list"].astype(np.int32), f["point_list"])
    assert np.array_equal(ours["ranges"].astype(np.int32), f["ranges"])
    # images and gradients: to rounding
    assert rel(ours["color"], f["color"]) <= 2e-5
    assert rel(ours["allmap"], f["allmap"]) <= 2e-5
    for k in ("means3D", "opacity", "scales", "rotations", "sh"):
        assert rel(ours["grads"][k], gr[k]) <= 2e-4, k
    assert rel(ours["grads"]["means2D"][:, :2], gr["means2D"][:, :2]) <= 2e-4
    assert np.isfinite(ours["grads"]["means3D"]).all()


@pytest.mark.parametrize("degree", [0, 1, 2])
def test_emulated_lower_sh_degrees(emu, degree):
    scene = synth.make_point_scene(250, seed=21, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    cam = synth.make_cameras(1, 32, 32, seed=22, device="cpu")[0]
    g = synth.upstream_grads(32, 32, 6, device="cpu")
    ours = run_emulated(emu, scene, cam, g["color"], g["allmap"], degree=degree)
    f = cpu_oracle.forward_scene(scene, cam, keep_state=True, sh_degree=degree)
    gr = cpu_oracle.backward(f, g["color"], g["allmap"])
    assert ours["R"] == f["num_rendered"] and ours["R"] > 30
    assert rel(ours["color"], f["color"]) <= 2e-5
    assert rel(ours["grads"]["sh"], gr["sh"]) <= 2e-4
    n_active = (degree + 1) ** 2
    assert not ours["grads"]["sh"][:, n_active:].any()      # coefficients above the active degree get zero gradient
    assert rel(ours["grads"]["means3D"], gr["means3D"]) <= 2e-4


def test_emulated_precomputed_colours(emu):
    scene = synth.make_point_scene(250, seed=31, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    cam = synth.make_cameras(1, 32, 32, seed=32, device="cpu")[0]
    g = synth.upstream_grads(32, 32, 7, device="cpu")
    colors = torch.rand(250, 3, generator=torch.Generator().manual_seed(1))
    ours = run_emulated(emu, scene, cam, g["color"], g["allmap"], colors_precomp=colors)
    # exactly one of SHs / precomputed colours, as the reference API enforces
    f = cpu_oracle.forward(scene["means3D"], scene["scales"], scene["rotations"], scene["opacities"], None,
                           cam.viewmatrix, cam.projmatrix, cam.campos, 32, 32, cam.tanfovx, cam.tanfovy,
                           colors_precomp=colors, keep_state=True)
    gr = cpu_oracle.backward(f, g["color"], g["allmap"])
    assert ours["R"] == f["num_rendered"] and ours["R"] > 30
    assert rel(ours["color"], f["color"]) <= 2e-5 and rel(ours["allmap"], f["allmap"]) <= 2e-5
    assert rel(ours["grads"]["colors"], gr["colors"]) <= 2e-4
    assert rel(ours["grads"]["means3D"], gr["means3D"]) <= 2e-4


def test_emulated_nothing_visible_and_tiny_image(emu):
    scene = synth.make_point_scene(100, seed=41, device="cpu")
    g = synth.upstream_grads(5, 3, 8, device="cpu")
    cam = synth.make_cameras(1, 5, 3, seed=42, device="cpu")[0]
    # everything behind the camera: no instance, background image, zero gradients
    far = dict(scene)
    far["means3D"] = scene["means3D"] + cam.campos * 3.0
    ours = run_emulated(emu, far, cam, g["color"], g["allmap"])
    assert ours["R"] == 0 and not ours["radii"].any()
    assert not ours["color"].any() and not ours["allmap"].any()
    for k in ("means3D", "opacity", "scales", "rotations", "sh"):
        assert not ours["grads"][k].any(), k
    # a 5x3 image (one ragged tile) with splats covering it
    scene["scales"] = scene["scales"] * 30.0
    ours = run_emulated(emu, scene, cam, g["color"], g["allmap"])
    f = cpu_oracle.forward_scene(scene, cam, keep_state=True)
    gr = cpu_oracle.backward(f, g["color"], g["allmap"])
    assert ours["R"] == f["num_rendered"]
    assert np.array_equal(ours["radii"], f["radii"])
    if ours["R"]:
        assert rel(ours["color"], f["color"]) <= 2e-5
        assert rel(ours["grads"]["means3D"], gr["means3D"]) <= 2e-4


@pytest.mark.parametrize("P,degree,M", [(517, 2, 16), (97, 1, 4), (64, 0, 1), (1, 3, 16)])
def test_emulated_cooperative_sh_rows(emu, P, degree, M):
    """The forward preprocess stages the SH rows warp-cooperatively through shared memory: ragged warps (P % 32 != 0), a
    single surfel, lower active degrees and narrower coefficient rows (M = 1, 4) against the C oracle."""
    scene = synth.make_point_scene(P, seed=31, device="cpu")
    scene["scales"] = scene["scales"] * 3.0
    scene["shs"] = scene["shs"][:, :M].contiguous()
    cam = synth.make_cameras(1, 40, 24, seed=32, device="cpu")[0]
    g = synth.upstream_grads(40, 24, 7, device="cpu")
    ours = run_emulated(emu, scene, cam, g["color"], g["allmap"], degree=degree)
    f = cpu_oracle.forward_scene(scene, cam, sh_degree=degree, keep_state=True)
    assert np.array_equal(ours["radii"], f["radii"])
    assert (ours["radii"] > 0).sum() >= min(P, 20) // 2
    assert rel(ours["color"], f["color"]) <= 2e-5

def test_emulated_binning_across_many_ctas_equals_the_reference_key_order(emu):
    """Production binning (depth sort of the surfels, depth-ordered emission, tile-id sort) with several CTAs in
    every look-back chain (3 emission CTAs, > 4 onesweep tiles): the instance list equals a stable sort of the
    reference's (tile << 32 | depth bits) keys in duplication order (rasterizer_impl.cu:70-111, 301-309), and the
    rebuilt 64-bit keys / tile ranges follow."""
    P, W, H = 2500, 96, 64
    scene = synth.make_point_scene(P, seed=77, device="cpu")
    scene["scales"] = scene["scales"] * 6.0
    cam = synth.make_cameras(1, W, H, seed=78, device="cpu")[0]
    g = synth.upstream_grads(W, H, 5, device="cpu")
    ours = run_emulated(emu, scene, cam, g["color"], g["allmap"])
    f = cpu_oracle.forward_scene(scene, cam, keep_state=True)
    assert np.array_equal(ours["radii"], f["radii"])
    R = ours["R"]
    assert R == f["num_rendered"] and R > 4 * 3072, R
    keys = ours["keys"]
    assert np.all(keys[1:] >= keys[:-1]), "sorted keys are not sorted"
    # stability: equal keys keep duplication order = ascending surfel index
    same = keys[1:] == keys[:-1]
    assert np.all(ours["point_list"][1:][same] > ours["point_list"][:-1][same])
    # ranges partition the list by tile
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        lo, hi = ours["ranges"][t]
        assert np.all(tiles[lo:hi] == t) and hi - lo == int((tiles == t).sum())
    # every surfel appears tiles_touched times
    counts = np.bincount(ours["point_list"], minlength=P)
    assert counts.sum() == R and np.array_equal(counts > 0, f["radii"] > 0)
