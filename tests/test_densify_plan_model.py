"""CPU model of the *algorithm* of csrc/densify.cu (classify once -> scan -> row map -> one gather -> children), in
numpy float32 with the kernels' index arithmetic, against the golden vectors of the unmodified reference.  It checks
the design — that a single planned compaction reproduces the reference's clone -> split -> prune row order, which
rows get zeroed moments, which normal draw each child consumes — without a GPU; the CUDA kernels themselves are
compared with the oracle in tests/test_gpu_zz_densify.py."""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = sorted((Path(__file__).parent / "golden").glob("densify_*.npz"))
NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
KEEP, CLONE, SPLIT, CHILD = 1, 2, 4, 8
f32 = np.float32


def classify(z, N):
    accum, denom = z["in_accum"].reshape(-1), z["in_denom"].reshape(-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        g = (accum / denom).astype(f32)
    g[np.isnan(g)] = 0
    s = np.exp(z["in_scaling"]).astype(f32)
    smax = s.max(axis=1)
    max_grad, thr = f32(z["max_grad"]), f32(float(z["percent_dense"]) * float(z["extent"]))
    min_op, ws_thr = f32(z["min_opacity"]), f32(0.1 * float(z["extent"]))
    use_ws = float(z["max_screen_size"]) > 0
    clone_sel = (np.sqrt(g * g) >= max_grad) & (smax <= thr)
    split_sel = (g >= max_grad) & (smax > thr)
    sig = (f32(1) / (f32(1) + np.exp(-z["in_opacity"].reshape(-1)).astype(f32))).astype(f32)
    transparent = sig < min_op
    pruned = transparent | (use_ws & (smax > ws_thr))
    inv = f32(1) / f32(0.8 * N)
    child = np.exp(np.log((s * inv).astype(f32)).astype(f32)).astype(f32)
    child_pruned = transparent | (use_ws & (child.max(axis=1) > ws_thr))
    code = np.zeros(len(g), np.uint8)
    code[~split_sel & ~pruned] |= KEEP
    code[clone_sel & ~pruned] |= CLONE
    code[split_sel] |= SPLIT
    code[split_sel & ~child_pruned] |= CHILD
    return code, inv


def plan(code, N):
    ex = lambda m: np.cumsum(m) - m  # exclusive scan
    flags = [(code & b) != 0 for b in (KEEP, CLONE, SPLIT, CHILD)]
    rank = [ex(f.astype(np.int64)) for f in flags]
    n_keep, n_clone, n_sel, n_child = (int(f.sum()) for f in flags)
    n_out = n_keep + n_clone + N * n_child
    src_row = np.full(n_out, -1, np.int64)
    sample_row = np.full(N * n_child, -1, np.int64)
    for i in range(len(code)):
        if flags[0][i]:
            src_row[rank[0][i]] = i
        if flags[1][i]:
            src_row[n_keep + rank[1][i]] = i
        if flags[3][i]:
            for r in range(N):
                c = r * n_child + rank[3][i]
                src_row[n_keep + n_clone + c] = i
                sample_row[c] = r * n_sel + rank[2][i]
    assert (src_row >= 0).all() and (sample_row >= 0).all()
    return src_row, sample_row, n_keep, n_clone, n_sel, n_child


@pytest.mark.parametrize("path", GOLDEN, ids=[p.stem for p in GOLDEN])
def test_single_pass_plan_reproduces_reference(path):
    z = np.load(path)
    N = 2
    code, inv = classify(z, N)
    src_row, sample_row, n_keep, n_clone, n_sel, n_child = plan(code, N)
    assert N * n_sel == z["z"].shape[0]
    assert len(src_row) == z["out_xyz"].shape[0]
    first_child = n_keep + n_clone
    for k in NAMES:
        got = z["in_" + k][src_row]
        if k in ("xyz", "scaling"):  # children are recomputed
            assert np.array_equal(got[:first_child], z["out_" + k][:first_child]), k
        else:
            assert np.array_equal(got, z["out_" + k]), k
        for mv in ("m_", "v_"):
            mom = z["in_" + mv + k][src_row].copy()
            mom[n_keep:] = 0  # zero_new
            assert np.array_equal(mom, z["out_" + mv + k]), (mv, k)
    assert np.array_equal(z["in_semantic"][src_row], z["out_semantic"])
    # children
    par = src_row[first_child:]
    zz = z["z"][sample_row]
    s = np.exp(z["in_scaling"][par]).astype(f32)
    a = np.stack([zz[:, 0] * s[:, 0], zz[:, 1] * s[:, 1], zz[:, 2] * 0], axis=1).astype(f32)
    q = z["in_rotation"][par]
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    r, x, y, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + w * w), 2 * (x * y - r * w), 2 * (x * w + r * y),
                  2 * (x * y + r * w), 1 - 2 * (x * x + w * w), 2 * (y * w - r * x),
                  2 * (x * w - r * y), 2 * (y * w + r * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
    xyz = np.einsum("nij,nj->ni", R, a) + z["in_xyz"][par]
    np.testing.assert_allclose(xyz, z["out_xyz"][first_child:], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.log(s * inv), z["out_scaling"][first_child:], rtol=1e-6, atol=1e-6)
