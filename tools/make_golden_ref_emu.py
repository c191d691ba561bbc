"""Golden vectors of the `_part` rasteriser from the UNMODIFIED reference CUDA source
(/root/reference/submodules/diff-surfel-rasterization_part/cuda_rasterizer/*.cu), executed on the CPU by the
lock-step emulator (tests/cuda_emu: CUDA threads as OS threads; CUB / cooperative groups / GLM replaced by host
stand-ins) — forward and backward of CudaRasterizer::Rasterizer on small synthetic scenes.  Build container only
(needs /root/reference); the fixtures are committed under tests/golden/ and let the GPU-less suite compare the
emulated product kernels with the reference's own code for the fork that has no C restatement.

    python tools/make_golden_ref_emu.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "cuda_emu"))
import build as emu_build  # noqa: E402

from partgs_b200 import synth  # noqa: E402


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


def make(lib, name, P, S, W, H, seed, degree=3, scale_mul=3.0, scale_modifier=1.0):
    scene = synth.make_point_scene(P, seed=seed, S=S, device="cpu")
    cam = synth.make_cameras(1, W, H, seed=seed + 10, device="cpu")[0]
    g = synth.upstream_grads(W, H, seed + 20, n_aux=8, S=S, device="cpu")
    m3, sc, rot, op, sh, sem = (f32(scene[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs", "semantics"))
    sc = (sc * scale_mul).astype(np.float32)
    vm, pm, cp = f32(cam.viewmatrix), f32(cam.projmatrix), f32(cam.campos)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    M = sh.shape[1]
    color = np.zeros((3, H, W), np.float32); semantic = np.zeros((S, H, W), np.float32)
    allmap = np.zeros((8, H, W), np.float32); radii = np.zeros(P, np.int32)
    lib.ref_part_forward.restype = C.c_int
    R = lib.ref_part_forward(P, degree, M, _p(bg), W, H, S, _p(m3), _p(sh), None, _p(op), _p(sem), _p(sc),
                             C.c_float(scale_modifier), _p(rot), _p(vm), _p(pm), _p(cp), C.c_float(cam.tanfovx),
                             C.c_float(cam.tanfovy), _p(color), _p(semantic), _p(allmap), _p(radii))
    gc, ga, gs = f32(g["color"]), f32(g["allmap"]), f32(g["semantic"])
    d = dict(means2D=np.zeros((P, 3), np.float32), normal=np.zeros((P, 3), np.float32),
             opacity=np.zeros((P, 1), np.float32), colors=np.zeros((P, 3), np.float32),
             semantics=np.zeros((P, S), np.float32), means3D=np.zeros((P, 3), np.float32),
             transMat=np.zeros((P, 9), np.float32), sh=np.zeros((P, M, 3), np.float32),
             scales=np.zeros((P, 2), np.float32), rotations=np.zeros((P, 4), np.float32))
    lib.ref_part_backward(P, degree, M, R, _p(bg), W, H, S, _p(m3), _p(sh), None, _p(sem), _p(sc), C.c_float(scale_modifier),
                          _p(rot), _p(vm), _p(pm), _p(cp), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(radii),
                          _p(gc), _p(gs), _p(ga), _p(d["means2D"]), _p(d["normal"]), _p(d["opacity"]), _p(d["colors"]),
                          _p(d["semantics"]), _p(d["means3D"]), _p(d["transMat"]), _p(d["sh"]), _p(d["scales"]),
                          _p(d["rotations"]))
    out = dict(means3D=m3, scales=sc, rotations=rot, opacities=op, shs=sh, semantics=sem, viewmatrix=vm, projmatrix=pm,
               campos=cp, bg=bg, scale_modifier=np.float64(scale_modifier), tanfov=np.array([cam.tanfovx, cam.tanfovy], np.float64), W=W, H=H, degree=degree,
               g_color=gc, g_allmap=ga, g_semantic=gs, R=R, color=color, semantic=semantic, allmap=allmap, radii=radii)
    out.update({"d_" + k: v for k, v in d.items() if k != "normal"})
    path = ROOT / "tests" / "golden" / f"ref_emu_part_{name}.npz"
    np.savez_compressed(path, **out)
    print(path, "R", R, "visible", int((radii > 0).sum()), "of", P, "finite grads",
          all(np.isfinite(v).all() for v in d.values()))


def make_base(lib, name, P, W, H, seed, degree=3, scale_mul=3.0, precomp=False, scale_modifier=1.0):
    scene = synth.make_point_scene(P, seed=seed, device="cpu")
    cam = synth.make_cameras(1, W, H, seed=seed + 10, device="cpu")[0]
    g = synth.upstream_grads(W, H, seed + 20, device="cpu")
    m3, sc, rot, op, sh = (f32(scene[k]) for k in ("means3D", "scales", "rotations", "opacities", "shs"))
    sc = (sc * scale_mul).astype(np.float32)
    vm, pm, cp = f32(cam.viewmatrix), f32(cam.projmatrix), f32(cam.campos)
    bg = np.array([0.3, 0.1, 0.2], np.float32)
    M = sh.shape[1]
    color = np.zeros((3, H, W), np.float32); allmap = np.zeros((7, H, W), np.float32); radii = np.zeros(P, np.int32)
    lib.ref_base_forward.restype = C.c_int
    T = None
    if precomp:
        # the per-surfel 3x3 ray-splat transform the reference itself computes, read back from a regular run of the
        # pinned C oracle (geometry state `transMat`), then handed in as transMat_precomp with no scales / rotations
        from oracle import cpu_oracle
        T = np.ascontiguousarray(cpu_oracle.forward(m3, sc, rot, op, sh, vm, pm, cp, W, H, cam.tanfovx, cam.tanfovy,
                                                    bg=bg, sh_degree=degree)["transMat"], dtype=np.float32)
    sc_in, rot_in = (None, None) if precomp else (sc, rot)
    R = lib.ref_base_forward(P, degree, M, _p(bg), W, H, _p(m3), _p(sh), None, _p(op), _p(sc_in), C.c_float(scale_modifier),
                             _p(rot_in), _p(T), _p(vm), _p(pm), _p(cp), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(color),
                             _p(allmap), _p(radii))
    gc, ga = f32(g["color"]), f32(g["allmap"])
    d = dict(means2D=np.zeros((P, 3), np.float32), normal=np.zeros((P, 3), np.float32),
             opacity=np.zeros((P, 1), np.float32), colors=np.zeros((P, 3), np.float32),
             means3D=np.zeros((P, 3), np.float32), transMat=np.zeros((P, 9), np.float32),
             sh=np.zeros((P, M, 3), np.float32), scales=np.zeros((P, 2), np.float32),
             rotations=np.zeros((P, 4), np.float32))
    lib.ref_base_backward(P, degree, M, R, _p(bg), W, H, _p(m3), _p(sh), None, _p(sc_in), C.c_float(scale_modifier), _p(rot_in),
                          _p(T), _p(vm), _p(pm), _p(cp), C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(radii), _p(gc), _p(ga),
                          _p(d["means2D"]), _p(d["normal"]), _p(d["opacity"]), _p(d["colors"]), _p(d["means3D"]),
                          _p(d["transMat"]), _p(d["sh"]), _p(d["scales"]), _p(d["rotations"]))
    out = dict(means3D=m3, scales=sc, rotations=rot, opacities=op, shs=sh, viewmatrix=vm, projmatrix=pm, campos=cp,
               bg=bg, scale_modifier=np.float64(scale_modifier), tanfov=np.array([cam.tanfovx, cam.tanfovy], np.float64), W=W, H=H, degree=degree, g_color=gc,
               g_allmap=ga, R=R, color=color, allmap=allmap, radii=radii)
    out.update({"d_" + k: v for k, v in d.items() if k != "normal"})
    if precomp:
        out["transMat_precomp"] = T
    path = ROOT / "tests" / "golden" / f"ref_emu_base_{name}.npz"
    np.savez_compressed(path, **out)
    print(path, "R", R, "visible", int((radii > 0).sum()), "of", P, "finite grads",
          all(np.isfinite(v).all() for v in d.values()))


def make_knn(lib):
    """distCUDA2 = SimpleKNN::knn of the reference, on point sets that exercise its box pruning: uniform, tightly
    clustered + outliers, coincident points, and fewer than four points (FLT_MAX placeholders stay in the mean)."""
    g = np.random.default_rng(7)
    sets = {"uniform_1500": g.normal(size=(1500, 3)),
            "clustered_1100": np.concatenate([g.normal(size=(1000, 3)) * 0.01 + 3.0, g.normal(size=(100, 3)) * 5.0]),
            "duplicates_40": np.repeat(g.normal(size=(10, 3)), 4, axis=0),
            "p3": g.normal(size=(3, 3)), "p2": g.normal(size=(2, 3)), "p1": g.normal(size=(1, 3)),
            "p1025": g.random(size=(1025, 3))}
    out = {}
    for name, pts in sets.items():
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        d = np.full(pts.shape[0], np.nan, np.float32)
        lib.ref_knn(pts.shape[0], _p(pts), _p(d))
        out[name + "_points"], out[name + "_dist2"] = pts, d
    path = ROOT / "tests" / "golden" / "ref_emu_knn.npz"
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape, float(np.nanmax(v))) for k, v in out.items() if k.endswith("dist2")})


if __name__ == "__main__":
    make_knn(C.CDLL(str(emu_build.build_reference("knn"))))
    base = C.CDLL(str(emu_build.build_reference("base")))
    make_base(base, "p300_48x32", P=300, W=48, H=32, seed=61)
    make_base(base, "p400_40x24_deg2", P=400, W=40, H=24, seed=62, degree=2)
    make_base(base, "precompT_p300_48x32", P=300, W=48, H=32, seed=63, precomp=True)
    make_base(base, "scalemod_p300_48x32", P=300, W=48, H=32, seed=64, scale_modifier=0.7, scale_mul=4.0)
    lib = C.CDLL(str(emu_build.build_reference("part")))
    make(lib, "p300_s5_48x32", P=300, S=5, W=48, H=32, seed=51)
    make(lib, "p400_s16_40x24_deg1", P=400, S=16, W=40, H=24, seed=52, degree=1)
    make(lib, "scalemod_p300_s3_48x32", P=300, S=3, W=48, H=32, seed=53, scale_modifier=0.7)
