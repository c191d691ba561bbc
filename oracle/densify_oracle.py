"""TEST INFRASTRUCTURE — torch restatement of PartGS's densification.

Only tests/ may import this module; the product path (partgs_b200/densify.py -> csrc/densify.cu) never does.

Restates, as ONE function on plain tensors (any device), what the part model does at a densification step:
``TwoGaussianModel.densify_and_prune`` = clone -> split -> prune including the optimiser-state surgery
(games/block_mesh_splatting/scene/two_gaussian_model.py:341-423 for prune_points / densification_postfix /
densify_and_split / densify_and_clone; scene/gaussian_model.py:384-436 for _prune_optimizer /
cat_tensors_to_optimizer; :495-509 for densify_and_prune itself).  The only change of interface: the standard-normal
draws behind ``torch.normal(mean=0, std=stds)`` (:388) are an argument ``z [N*Ns,3]`` (ATen evaluates a tensor-tensor
normal as ``out.normal_(0,1).mul_(std).add_(mean)``), so the restatement is deterministic and can be compared with the
CUDA path on identical draws.

Pinned by golden vectors produced by the UNMODIFIED reference classes run on CPU (tools/make_golden_densify.py ->
tests/golden/densify_*.npz, tests/test_densify_oracle.py).
"""
from __future__ import annotations

import torch

PARAM_NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")  # optimiser group names, training_setup order


def build_rotation(r):
    """utils/general_utils.py:149-170."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r_ * z)
    R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y)
    R[:, 2, 1] = 2 * (y * z + r_ * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def split_selection(xyz_gradient_accum, denom, scaling, max_grad, extent, percent_dense):
    """(clone mask, split mask) over the P rows that exist before densification; the split mask is what sizes the
    reference's normal draw: Ns = split.sum(), z has N*Ns rows."""
    grads = xyz_gradient_accum / denom
    grads[grads.isnan()] = 0.0
    big = torch.max(torch.exp(scaling), dim=1).values > percent_dense * extent
    clone = (torch.norm(grads, dim=-1) >= max_grad) & ~big
    split = (grads.squeeze(-1) >= max_grad) & big
    return clone, split


def densify_and_prune(params, moments, semantic, xyz_gradient_accum, denom, max_grad, min_opacity, extent,
                      max_screen_size, percent_dense, z, N=2):
    """params: {name: tensor [P,...]} for PARAM_NAMES (raw, pre-activation); moments: {name: (exp_avg, exp_avg_sq)} or
    {name: None} for groups without optimiser state; semantic [P,S]; z [N*Ns,3] standard-normal draws.
    Returns (new params, new moments, new semantic, info) with the reference's row order:
    [surviving originals | surviving clones | surviving split children, replica-major]."""
    P = params["xyz"].shape[0]
    dev = params["xyz"].device
    cur = {k: v.detach().clone() for k, v in params.items()}
    mom = {k: (None if moments.get(k) is None else tuple(t.clone() for t in moments[k])) for k in PARAM_NAMES}
    sem = semantic.clone()

    def cat(new, new_sem):
        nonlocal sem
        for k in PARAM_NAMES:
            if mom[k] is not None:
                mom[k] = tuple(torch.cat((t, torch.zeros_like(new[k])), dim=0) for t in mom[k])
            cur[k] = torch.cat((cur[k], new[k]), dim=0)
        sem = torch.cat((sem, new_sem), dim=0)

    def prune(mask):
        nonlocal sem
        valid = ~mask
        for k in PARAM_NAMES:
            if mom[k] is not None:
                mom[k] = tuple(t[valid] for t in mom[k])
            cur[k] = cur[k][valid]
        sem = sem[valid]

    grads = xyz_gradient_accum / denom
    grads[grads.isnan()] = 0.0
    thr = percent_dense * extent

    # densify_and_clone
    sel = torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(torch.exp(cur["scaling"]), dim=1).values <= thr)
    n_clone_sel = int(sel.sum())
    cat({k: cur[k][sel] for k in PARAM_NAMES}, sem[sel])

    # densify_and_split
    n_init = cur["xyz"].shape[0]
    padded = torch.zeros((n_init,), device=dev)
    padded[:P] = grads.squeeze()
    sel = torch.where(padded >= max_grad, True, False)
    sel = torch.logical_and(sel, torch.max(torch.exp(cur["scaling"]), dim=1).values > thr)
    n_split_sel = int(sel.sum())
    stds = torch.exp(cur["scaling"][sel]).repeat(N, 1)
    stds = torch.cat([stds, 0 * torch.ones_like(stds[:, :1])], dim=-1)
    assert z.shape == stds.shape, (z.shape, stds.shape)
    samples = z * stds
    rots = build_rotation(cur["rotation"][sel]).repeat(N, 1, 1)
    new = {
        "xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + cur["xyz"][sel].repeat(N, 1),
        "scaling": torch.log(torch.exp(cur["scaling"][sel]).repeat(N, 1) / (0.8 * N)),
        "rotation": cur["rotation"][sel].repeat(N, 1),
        "f_dc": cur["f_dc"][sel].repeat(N, 1, 1),
        "f_rest": cur["f_rest"][sel].repeat(N, 1, 1),
        "opacity": cur["opacity"][sel].repeat(N, 1),
    }
    cat(new, sem[sel].repeat(N, 1))
    prune(torch.cat((sel, torch.zeros(N * n_split_sel, device=dev, dtype=torch.bool))))

    # final prune (max_radii2D was zeroed by densification_postfix, so the screen-size test never fires)
    prune_mask = (torch.sigmoid(cur["opacity"]) < min_opacity).squeeze(-1)
    if max_screen_size:
        max_radii2D = torch.zeros((cur["xyz"].shape[0],), device=dev)
        big_vs = max_radii2D > max_screen_size
        big_ws = torch.exp(cur["scaling"]).max(dim=1).values > 0.1 * extent
        prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
    prune(prune_mask)
    info = {"n_clone_selected": n_clone_sel, "n_split_selected": n_split_sel, "n_out": cur["xyz"].shape[0]}
    return cur, mom, sem, info
