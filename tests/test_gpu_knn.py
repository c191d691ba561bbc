"""distCUDA2 parity: product grid kNN vs the reference simple-knn build (oracle/_ref) and an
exact brute-force (torch.cdist in float64) on small clouds.  Tolerance 1e-6 relative
(exact 3-NN)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def brute(points):
    d = torch.cdist(points.double(), points.double()) ** 2
    d.fill_diagonal_(float("inf"))
    return d.topk(3, dim=1, largest=False).values.mean(dim=1).float()


def clouds(gen):
    from partgs_b200 import synth
    surf = synth.make_point_scene(20_000, 11)["means3D"]
    vol = torch.rand(15_000, 3, generator=gen) * torch.tensor([4.0, 1.0, 0.25])
    clustered = torch.cat([torch.randn(5000, 3, generator=gen) * 0.01, torch.randn(5000, 3, generator=gen) * 0.01 + 5.0,
                           torch.tensor([[100.0, -50.0, 3.0]])])
    flat = torch.cat([torch.rand(8000, 2, generator=gen), torch.zeros(8000, 1)], dim=1)   # degenerate axis
    dup = torch.rand(3000, 3, generator=gen).repeat(2, 1)                                  # exact duplicates
    return dict(surface=surf, volume=vol, clustered=clustered, flat=flat, duplicates=dup)


def test_knn_vs_bruteforce_and_reference():
    from partgs_b200.simple_knn._C import distCUDA2
    from oracle import ref_cuda
    gen = torch.Generator().manual_seed(5)
    ref = ref_cuda.load("ref_knn_C") if ref_cuda.available("ref_knn_C") else None
    for name, pts in clouds(gen).items():
        pts = pts.to(DEV).contiguous()
        ours = distCUDA2(pts)
        exact = brute(pts)
        assert torch.allclose(ours, exact, rtol=1e-5, atol=1e-12), name
        if ref is not None:
            r = ref.distCUDA2(pts)
            torch.cuda.synchronize()
            assert torch.allclose(ours, r, rtol=1e-6, atol=1e-12), name


def test_knn_small_and_empty():
    from partgs_b200.simple_knn._C import distCUDA2
    assert distCUDA2(torch.zeros(0, 3, device=DEV)).shape == (0,)
    # fewer than 4 points: missing neighbours stay at FLT_MAX like the reference (simple_knn.cu:154,182)
    out = distCUDA2(torch.tensor([[0.0, 0, 0], [1.0, 0, 0]], device=DEV))
    assert bool((out > 1e37).all())
    pts = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [0, 2.0, 0], [0, 0, 3.0]], device=DEV)
    assert torch.allclose(distCUDA2(pts), brute(pts))
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(4, 3))          # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(4, 2, device=DEV))


def test_knn_full_size_properties():
    """1M / 3M points: finite, positive, invariant under permutation and translation."""
    from partgs_b200 import synth
    from partgs_b200.simple_knn._C import distCUDA2
    pts = synth.make_point_scene(1_000_000, 42, device=DEV)["means3D"]
    d = distCUDA2(pts)
    assert bool(torch.isfinite(d).all()) and float(d.min()) >= 0
    perm = torch.randperm(pts.shape[0], device=DEV)
    dp = distCUDA2(pts[perm].contiguous())
    assert torch.allclose(dp, d[perm], rtol=1e-6, atol=0)
    sub = torch.arange(0, pts.shape[0], 997, device=DEV)
    exact = (torch.cdist(pts[sub].double(), pts.double()) ** 2)
    exact[torch.arange(sub.numel()), sub] = float("inf")
    exact = exact.topk(3, dim=1, largest=False).values.mean(dim=1).float()
    assert torch.allclose(d[sub], exact, rtol=1e-5, atol=1e-14)
    from oracle import ref_cuda
    if ref_cuda.available("ref_knn_C"):
        r = ref_cuda.load("ref_knn_C").distCUDA2(pts)
        assert torch.allclose(d, r, rtol=1e-6, atol=1e-14)
