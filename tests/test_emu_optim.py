"""csrc/optim.cu (one-launch Adam over several tensors, densification statistics) compiled for the CPU lock-step
emulator (tests/cuda_emu) against torch's own CPU Adam and the reference's torch expressions — a GPU-less regression
test of kernels whose hardware parity is covered by tests/test_gpu_optim.py."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).parent / "cuda_emu"))
import build as emu_build  # noqa: E402

EXPORTS = r'''
extern "C" {
void emu_adam(int n, float* const* p, const float* const* g, float* const* m, float* const* v, const size_t* numel,
              const float* step_size, double beta1, double beta2, double eps, double bc2_sqrt) {
  pgs::AdamTable t;
  t.n = n;
  for (int i = 0; i < n; i++) {
    t.param[i] = p[i]; t.grad[i] = g[i]; t.exp_avg[i] = m[i]; t.exp_avg_sq[i] = v[i];
    t.numel[i] = numel[i]; t.step_size[i] = step_size[i];
  }
  pgs::launch_adam_multi(t, beta1, beta2, eps, bc2_sqrt, nullptr);
}
void emu_stats(int P, const int* radii, const float* g2d, float* max_radii, float* accum, float* denom) {
  pgs::launch_densify_stats(P, radii, g2d, max_radii, accum, denom, nullptr);
}
}
'''


@pytest.fixture(scope="module")
def emu():
    try:
        return C.CDLL(str(emu_build.build("optim.cu", EXPORTS)))
    except emu_build.EmuUnavailable as ex:
        pytest.skip(str(ex))


def test_emulated_adam_matches_torch(emu):
    gen = torch.Generator().manual_seed(0)
    shapes = [(700, 3), (700, 1, 3), (700, 3, 3), (700, 1), (1100,), (5,)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 0.05, 0.005, 0.0]
    ref_p = [torch.randn(s, generator=gen).requires_grad_(True) for s in shapes]
    ours_p = [p.detach().numpy().copy() for p in ref_p]
    ours_m = [np.zeros_like(p) for p in ours_p]
    ours_v = [np.zeros_like(p) for p in ours_p]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref_p, lrs)], lr=0.0, eps=1e-15, foreach=False)
    beta1, beta2, eps = 0.9, 0.999, 1e-15
    n = len(shapes)
    for step in range(1, 4):
        grads = [torch.randn(s, generator=gen) * 10.0 ** (step - 2) for s in shapes]
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        opt.step()
        gs = [g.numpy().copy() for g in grads]
        arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        emu.emu_adam(n, arr(ours_p), arr(gs), arr(ours_m), arr(ours_v), (C.c_size_t * n)(*[p.size for p in ours_p]),
                     (C.c_float * n)(*[lr / (1 - beta1 ** step) for lr in lrs]), C.c_double(beta1), C.c_double(beta2),
                     C.c_double(eps), C.c_double((1 - beta2 ** step) ** 0.5))
        for i, p in enumerate(ref_p):
            st = opt.state[p]
            np.testing.assert_allclose(ours_p[i], p.detach().numpy(), rtol=3e-6, atol=5e-7, err_msg=f"param {i} step {step}")
            np.testing.assert_allclose(ours_m[i], st["exp_avg"].numpy(), rtol=3e-6, atol=2e-6 * float(st["exp_avg"].abs().max()))
            np.testing.assert_allclose(ours_v[i], st["exp_avg_sq"].numpy(), rtol=3e-6, atol=2e-6 * float(st["exp_avg_sq"].abs().max()))
    assert np.array_equal(ours_p[-1], ref_p[-1].detach().numpy())  # lr 0: untouched


def test_emulated_densification_stats(emu):
    g = np.random.default_rng(3)
    P = 1000
    radii = (g.integers(-1, 30, P)).astype(np.int32)
    g2d = g.normal(size=(P, 3)).astype(np.float32)
    max_r = (g.random(P) * 20).astype(np.float32)
    accum = g.random((P, 1)).astype(np.float32)
    denom = g.integers(0, 5, (P, 1)).astype(np.float32)
    vis = radii > 0
    want_max = max_r.copy(); want_max[vis] = np.maximum(max_r[vis], radii[vis].astype(np.float32))
    want_acc = accum.copy(); want_acc[vis, 0] += np.sqrt(g2d[vis, 0] ** 2 + g2d[vis, 1] ** 2)
    want_den = denom.copy(); want_den[vis] += 1
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.emu_stats(P, p(radii), p(g2d), p(max_r), p(accum), p(denom))
    assert np.array_equal(max_r, want_max) and np.array_equal(denom, want_den)
    np.testing.assert_allclose(accum, want_acc, rtol=1e-6)
