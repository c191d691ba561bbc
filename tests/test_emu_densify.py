"""The ACTUAL kernel source of csrc/densify.cu and csrc/extract.cu, compiled for the CPU lock-step emulator
(tests/cuda_emu) and run on the golden vectors of the unmodified reference: indexing, ballots, the count scan,
the row map, the multi-tensor gather and the children arithmetic are exercised without a GPU.  (Transcendentals come
from glibc here, so decisions exactly on a threshold could differ from the GPU's; the fixtures have none.)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "cuda_emu"))
import build as emu_build  # noqa: E402

GOLDEN = sorted((Path(__file__).parent / "golden").glob("densify_*.npz"))
NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")

EXPORTS = r'''
extern "C" {
int emu_blocks(int P) { return pgs::densify_blocks(P); }
void emu_plan(int P, const float* accum, const float* denom, const float* scaling, const float* opacity,
              float max_grad, float dense_thr, float min_opacity, int use_ws, float ws_thr, float inv_divisor,
              unsigned char* code, uint32_t* block_off, uint32_t* counts) {
  pgs::launch_densify_plan(P, accum, denom, scaling, opacity, max_grad, dense_thr, min_opacity, use_ws, ws_thr,
                           inv_divisor, code, block_off, counts, nullptr);
}
void emu_map(int P, const unsigned char* code, const uint32_t* block_off, const uint32_t* counts, int n_split,
             int* src_row, int* sample_row) {
  pgs::launch_densify_map(P, code, block_off, counts, n_split, src_row, sample_row, nullptr);
}
void emu_gather(int n, const float* const* src, float* const* dst, const int* widths, const int* zero_new, int n_out,
                int n_keep, const int* src_row) {
  pgs::GatherTable t;
  t.n = n;
  for (int i = 0; i < n; i++) {
    t.src[i] = src[i]; t.dst[i] = dst[i]; t.width[i] = widths[i]; t.zero_new[i] = zero_new[i];
    t.numel[i] = (size_t)n_out * widths[i];
  }
  pgs::launch_densify_gather(t, n_keep, src_row, nullptr);
}
void emu_children(int n_children, const uint32_t* counts, const int* src_row, const int* sample_row, const float* z,
                  const float* xyz_in, const float* scaling_in, const float* rotation_in, float inv_divisor,
                  float* xyz_out, float* scaling_out) {
  pgs::launch_densify_children(n_children, counts, src_row, sample_row, z, xyz_in, scaling_in, rotation_in,
                               inv_divisor, xyz_out, scaling_out, nullptr);
}
}
'''


@pytest.fixture(scope="module")
def emu():
    try:
        return C.CDLL(str(emu_build.build("densify.cu", EXPORTS)))
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run_emulated(emu, z, N=2):
    f32 = np.float32
    P = z["in_xyz"].shape[0]
    c = lambda a: np.ascontiguousarray(a, dtype=f32)
    accum, denom, scaling, opacity = c(z["in_accum"]), c(z["in_denom"]), c(z["in_scaling"]), c(z["in_opacity"])
    nblk = emu.emu_blocks(P)
    code = np.zeros(P, np.uint8)
    block_off = np.zeros(4 * nblk, np.uint32)
    counts = np.full(8, 0xdeadbeef, np.uint32)
    extent = float(z["extent"])
    inv = f32(1) / f32(0.8 * N)
    emu.emu_plan(P, _p(accum), _p(denom), _p(scaling), _p(opacity), C.c_float(float(z["max_grad"])),
                 C.c_float(float(z["percent_dense"]) * extent), C.c_float(float(z["min_opacity"])),
                 1 if float(z["max_screen_size"]) > 0 else 0, C.c_float(0.1 * extent), C.c_float(inv), _p(code),
                 _p(block_off), _p(counts))
    n_keep, n_clone, n_sel, n_child = (int(v) for v in counts[:4])
    n_out = n_keep + n_clone + N * n_child
    src_row = np.full(max(n_out, 1), -1, np.int32)
    sample_row = np.full(max(N * n_child, 1), -1, np.int32)
    emu.emu_map(P, _p(code), _p(block_off), _p(counts), N, _p(src_row), _p(sample_row))
    srcs, dsts, widths, zero_new, out = [], [], [], [], {}
    for k in NAMES:
        for pre, zn in (("", 0), ("m_", 1), ("v_", 1)):
            a = c(z["in_" + pre + k])
            w = int(np.prod(a.shape[1:]))
            d = np.full((n_out,) + a.shape[1:], np.nan, f32)
            out[pre + k] = d
            if w:
                srcs.append(a); dsts.append(d); widths.append(w); zero_new.append(zn)
    sem = c(z["in_semantic"])
    out["semantic"] = np.full((n_out,) + sem.shape[1:], np.nan, f32)
    srcs.append(sem); dsts.append(out["semantic"]); widths.append(sem.shape[1]); zero_new.append(0)
    n = len(srcs)
    emu.emu_gather(n, (C.c_void_p * n)(*[a.ctypes.data for a in srcs]), (C.c_void_p * n)(*[a.ctypes.data for a in dsts]),
                   (C.c_int * n)(*widths), (C.c_int * n)(*zero_new), n_out, n_keep, _p(src_row))
    zz = c(z["z"])
    emu.emu_children(N * n_child, _p(counts), _p(src_row), _p(sample_row), _p(zz), _p(c(z["in_xyz"])), _p(scaling),
                     _p(c(z["in_rotation"])), C.c_float(inv), _p(out["xyz"]), _p(out["scaling"]))
    return out, dict(n_keep=n_keep, n_clone=n_clone, n_sel=n_sel, n_child=n_child, n_out=n_out, counts=counts)


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_emulated_kernels_reproduce_reference(emu, path):
    z = np.load(path)
    out, info = run_emulated(emu, z)
    assert info["n_out"] == z["out_xyz"].shape[0]
    assert 2 * info["n_sel"] == z["z"].shape[0]
    first_child = info["n_keep"] + info["n_clone"]
    for k in NAMES:
        if k in ("xyz", "scaling"):
            assert np.array_equal(out[k][:first_child], z["out_" + k][:first_child]), k
            np.testing.assert_allclose(out[k][first_child:], z["out_" + k][first_child:], rtol=2e-5, atol=2e-6)
        else:
            assert np.array_equal(out[k], z["out_" + k]), k
        assert np.array_equal(out["m_" + k], z["out_m_" + k]), k
        assert np.array_equal(out["v_" + k], z["out_v_" + k]), k
    assert np.array_equal(out["semantic"], z["out_semantic"])
    assert int(info["counts"][5]) == 0 and int(info["counts"][4]) >= info["n_clone"]


def test_emulated_scan_carries_across_chunks_of_1024_ctas(emu):
    """More than 1024 CTAs: the single-CTA scan of the per-CTA counts has to carry between its chunks."""
    g = np.random.default_rng(0)
    P = 256 * 1030 + 17
    z = {"in_xyz": np.zeros((P, 3), np.float32), "in_accum": (g.random((P, 1)) * 6e-4).astype(np.float32),
         "in_denom": np.ones((P, 1), np.float32), "in_scaling": np.log(10 ** (g.random((P, 2)) * 3 - 3.2)).astype(np.float32),
         "in_opacity": (g.normal(size=(P, 1)) * 3 - 2).astype(np.float32), "extent": 1.0, "max_grad": 0.0002,
         "percent_dense": 0.01, "min_opacity": 0.005, "max_screen_size": 20.0}
    f32 = np.float32
    nblk = emu.emu_blocks(P)
    assert nblk > 1024
    code = np.zeros(P, np.uint8); block_off = np.zeros(4 * nblk, np.uint32); counts = np.zeros(8, np.uint32)
    emu.emu_plan(P, _p(z["in_accum"]), _p(z["in_denom"]), _p(z["in_scaling"]), _p(z["in_opacity"]), C.c_float(0.0002),
                 C.c_float(0.01), C.c_float(0.005), 1, C.c_float(0.1), C.c_float(f32(1) / f32(1.6)), _p(code),
                 _p(block_off), _p(counts))
    for row, bit in enumerate((1, 2, 4, 8)):
        per_cta = np.add.reduceat(((code & bit) != 0).astype(np.int64), np.arange(0, P, 256))
        want = np.cumsum(per_cta) - per_cta
        assert np.array_equal(block_off[row * nblk:(row + 1) * nblk].astype(np.int64), want), row
        assert int(counts[row]) == int(per_cta.sum())
    assert int(counts[4]) == int(((code & 16) != 0).sum())
    n_keep, n_clone, n_sel, n_child = (int(v) for v in counts[:4])
    n_out = n_keep + n_clone + 2 * n_child
    src_row = np.full(n_out, -1, np.int32); sample_row = np.full(2 * n_child, -1, np.int32)
    emu.emu_map(P, _p(code), _p(block_off), _p(counts), 2, _p(src_row), _p(sample_row))
    assert (src_row >= 0).all() and (sample_row >= 0).all()
    assert np.array_equal(src_row[:n_keep], np.nonzero(code & 1)[0])
    assert np.array_equal(src_row[n_keep:n_keep + n_clone], np.nonzero(code & 2)[0])
    kids = np.nonzero(code & 8)[0]
    assert np.array_equal(src_row[n_keep + n_clone:], np.concatenate([kids, kids]))
    sel_rank = np.cumsum((code & 4) != 0) - 1
    assert np.array_equal(sample_row, np.concatenate([sel_rank[kids], n_sel + sel_rank[kids]]))
