"""Densification (clone / split / prune) of the point-level surfel model — SURVEY.md §8(f) rank 3.

Drop-in for ``TwoGaussianModel.densify_and_prune`` (games/block_mesh_splatting/scene/two_gaussian_model.py:341-423;
scene/gaussian_model.py:384-436, 495-509), including the optimiser-state surgery of ``_prune_optimizer`` /
``cat_tensors_to_optimizer``.  The reference moves the whole model and both Adam moments five times through boolean
masks and ``torch.cat``; here the surfels are classified once, the reference's final row order
``[surviving originals | surviving clones | surviving split children]`` is planned by a scan, and every tensor is
moved once by a single multi-tensor gather launch (csrc/densify.cu behind ``pgs_densify_*``).  One host read-back
(the counts that size the new tensors) instead of a dozen.  CUDA float32 tensors only; anything else raises.

The split draws ``torch.normal(0, stds)`` in the reference; ATen evaluates that as ``normal_(0,1) * std``, so the
same Philox draws are consumed here by ``torch.empty(N*Ns, 3).normal_()`` — a run seeded like the reference produces
the reference's children.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib

PARAM_NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")  # optimiser group names (training_setup)
_MODEL_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
               "scaling": "_scaling", "rotation": "_rotation"}


def _check(t: torch.Tensor, name: str, rows: int) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.dtype == torch.float32):
        raise RuntimeError(f"densify: {name} must be a CUDA float32 tensor (no fallback)")
    t = _lib.require_cuda_float(t.detach(), f"densify: {name}")  # raises for CPU tensors
    if t.shape[0] != rows:
        raise RuntimeError(f"densify: {name} has {t.shape[0]} rows, expected {rows}")
    return t


def densify_and_prune(params: Dict[str, torch.Tensor], moments: Dict[str, Optional[Tuple[torch.Tensor, torch.Tensor]]],
                      semantic: Optional[torch.Tensor], xyz_gradient_accum: torch.Tensor, denom: torch.Tensor,
                      max_grad: float, min_opacity: float, extent: float, max_screen_size, percent_dense: float,
                      N: int = 2, z: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
    """Functional core.  ``params``: raw (pre-activation) tensors for PARAM_NAMES, ``[P, ...]``; ``moments[name]``:
    ``(exp_avg, exp_avg_sq)`` or ``None`` (group without optimiser state); ``semantic [P,S]`` or ``None``.
    Returns ``(new_params, new_moments, new_semantic, info)``; ``info`` holds the counts.  ``xyz_gradient_accum`` /
    ``denom`` are only read (the caller replaces them by zeros of the new size, like densification_postfix)."""
    lib = _lib.load()
    P = params["xyz"].shape[0]
    dev = params["xyz"].device
    src = {k: _check(params[k], k, P) for k in PARAM_NAMES}
    if src["scaling"].numel() != 2 * P or src["rotation"].numel() != 4 * P or src["xyz"].numel() != 3 * P or \
            src["opacity"].numel() != P:
        raise RuntimeError("densify: expected xyz [P,3], opacity [P,1], scaling [P,2], rotation [P,4]")
    accum = _check(xyz_gradient_accum, "xyz_gradient_accum", P)
    den = _check(denom, "denom", P)
    if accum.numel() != P or den.numel() != P:
        raise RuntimeError("densify: xyz_gradient_accum / denom must hold one value per surfel")
    N = int(N)
    if N < 1:
        raise ValueError("densify: N must be >= 1")
    stream = _lib.current_stream(dev)
    nblk = lib.pgs_densify_blocks(P)
    code = torch.empty((max(P, 1),), dtype=torch.uint8, device=dev)
    block_off = torch.empty((max(4 * nblk, 1),), dtype=torch.int32, device=dev)
    counts = torch.empty((8,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgs_densify_plan(P, accum.data_ptr(), den.data_ptr(), src["scaling"].data_ptr(),
                                  src["opacity"].data_ptr(), float(max_grad), float(percent_dense * extent),
                                  float(min_opacity), 1 if max_screen_size else 0, float(0.1 * extent),
                                  float(0.8 * N), code.data_ptr(), block_off.data_ptr(), counts.data_ptr(), stream)
    _lib.check(rc, "pgs_densify_plan")
    n_keep, n_clone, n_sel, n_child, n_clone_sel = (int(v) for v in counts[:5].tolist())  # the one read-back
    n_out = n_keep + n_clone + N * n_child
    info = {"n_in": P, "n_out": n_out, "n_kept": n_keep, "n_clones": n_clone, "n_clone_selected": n_clone_sel,
            "n_split_selected": n_sel, "n_children": N * n_child}

    if z is None:
        z = torch.empty((N * n_sel, 3), dtype=torch.float32, device=dev).normal_(generator=generator)
    else:
        z = _check(z, "z", N * n_sel)
        if z.numel() != 3 * N * n_sel:
            raise RuntimeError("densify: z must be [N*Ns, 3]")

    src_row = torch.empty((max(n_out, 1),), dtype=torch.int32, device=dev)
    sample_row = torch.empty((max(N * n_child, 1),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgs_densify_map(P, code.data_ptr(), block_off.data_ptr(), counts.data_ptr(), N, src_row.data_ptr(),
                                 sample_row.data_ptr(), stream)
    _lib.check(rc, "pgs_densify_map")

    # one gather launch for every tensor that follows the surfels
    jobs = []  # (source, destination, width, zero_new)
    new_params, new_moments = {}, {}
    for k in PARAM_NAMES:
        s = src[k]
        new_params[k] = torch.empty((n_out,) + tuple(s.shape[1:]), dtype=torch.float32, device=dev)
        width = int(math.prod(s.shape[1:]))  # 0 for f_rest at SH degree 0: nothing to move
        if width:
            jobs.append((s, new_params[k], width, 0))
        mom = moments.get(k) if moments is not None else None
        if mom is None:
            new_moments[k] = None
        else:
            pair = []
            for j, m in enumerate(mom):
                m = _check(m, f"moment {j} of {k}", P)
                if m.shape != s.shape:
                    raise RuntimeError(f"densify: optimiser state of {k} does not match the parameter's shape")
                out = torch.empty_like(new_params[k])
                pair.append(out)
                if width:
                    jobs.append((m, out, width, 1))
            new_moments[k] = tuple(pair)
    new_semantic = None
    if semantic is not None:
        sem = _check(semantic, "semantic", P)
        new_semantic = torch.empty((n_out,) + tuple(sem.shape[1:]), dtype=torch.float32, device=dev)
        if sem.numel() > 0:
            jobs.append((sem, new_semantic, int(math.prod(sem.shape[1:])), 0))
    if n_out > 0 and P > 0:
        for s0 in range(0, len(jobs), 24):
            chunk = jobs[s0:s0 + 24]
            n = len(chunk)
            sa = (C.c_void_p * n)(*[j[0].data_ptr() for j in chunk])
            da = (C.c_void_p * n)(*[j[1].data_ptr() for j in chunk])
            wa = (C.c_int * n)(*[int(j[2]) for j in chunk])
            za = (C.c_int * n)(*[int(j[3]) for j in chunk])
            with torch.cuda.device(dev):
                rc = lib.pgs_densify_gather(n, sa, da, wa, za, n_out, n_keep, src_row.data_ptr(), stream)
            _lib.check(rc, "pgs_densify_gather")
        if N * n_child > 0:
            with torch.cuda.device(dev):
                rc = lib.pgs_densify_children(N * n_child, counts.data_ptr(), src_row.data_ptr(),
                                              sample_row.data_ptr(), z.data_ptr(), src["xyz"].data_ptr(),
                                              src["scaling"].data_ptr(), src["rotation"].data_ptr(), float(0.8 * N),
                                              new_params["xyz"].data_ptr(), new_params["scaling"].data_ptr(), stream)
            _lib.check(rc, "pgs_densify_children")
    return new_params, new_moments, new_semantic, info


def rewrap_optimizer(optimizer: torch.optim.Optimizer, new_params: Dict[str, torch.Tensor],
                     new_moments: Dict[str, Optional[Tuple[torch.Tensor, torch.Tensor]]]) -> Dict[str, nn.Parameter]:
    """What ``_prune_optimizer`` / ``cat_tensors_to_optimizer`` do to the optimiser (scene/gaussian_model.py:384-436):
    every named group gets a fresh ``nn.Parameter``; its state dict (``step`` kept, moments replaced) moves to the
    new key.  Groups whose name is not in ``new_params`` are left alone.  Device-agnostic host logic."""
    out = {}
    for group in optimizer.param_groups:
        name = group.get("name")
        if name not in new_params:
            continue
        if len(group["params"]) != 1:
            raise RuntimeError(f"optimiser group {name!r} must hold exactly one tensor (as in training_setup)")
        old = group["params"][0]
        stored = optimizer.state.get(old, None)
        new = nn.Parameter(new_params[name].requires_grad_(True))
        if stored is not None:
            mom = new_moments.get(name)
            if mom is None:
                raise RuntimeError(f"optimiser group {name!r} has state but no new moments were supplied")
            stored["exp_avg"], stored["exp_avg_sq"] = mom
            del optimizer.state[old]
            optimizer.state[new] = stored
        group["params"][0] = new
        out[name] = new
    return out


def densify_and_prune_model(model, max_grad, min_opacity, extent, max_screen_size, N: int = 2, generator=None):
    """``model.densify_and_prune(max_grad, min_opacity, extent, max_screen_size)`` for a reference-style model object
    (attributes ``_xyz, _features_dc, _features_rest, _opacity, _scaling, _rotation, _semantic, xyz_gradient_accum,
    denom, max_radii2D, percent_dense, optimizer`` with the groups of ``training_setup``).  Returns the counts."""
    opt = model.optimizer
    by_name = {g["name"]: g["params"][0] for g in opt.param_groups if g.get("name") in PARAM_NAMES}
    missing = [k for k in PARAM_NAMES if k not in by_name]
    if missing:
        raise RuntimeError(f"optimiser lacks the parameter groups {missing}")
    params = {k: by_name[k] for k in PARAM_NAMES}
    moments = {}
    for k in PARAM_NAMES:
        st = opt.state.get(by_name[k], None)
        moments[k] = None if not st else (st["exp_avg"], st["exp_avg_sq"])
    semantic = getattr(model, "_semantic", None)
    new_params, new_moments, new_semantic, info = densify_and_prune(
        params, moments, semantic, model.xyz_gradient_accum, model.denom, max_grad, min_opacity, extent,
        max_screen_size, model.percent_dense, N=N, generator=generator)
    wrapped = rewrap_optimizer(opt, new_params, new_moments)
    for k, attr in _MODEL_ATTR.items():
        setattr(model, attr, wrapped[k])
    if semantic is not None:
        model._semantic = new_semantic
    dev = new_params["xyz"].device
    n = info["n_out"]
    model.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
    model.denom = torch.zeros((n, 1), device=dev)
    model.max_radii2D = torch.zeros((n,), device=dev)
    return info
