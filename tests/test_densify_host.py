"""Host-side logic of partgs_b200.densify that needs no GPU: the optimiser re-wrap (what the reference's
_prune_optimizer / cat_tensors_to_optimizer do, scene/gaussian_model.py:384-436) and the loud failures."""
import pytest
import torch

from oracle import densify_oracle
from partgs_b200 import densify


def _optimizer(P=50, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 3, 3), "opacity": (P, 1), "scaling": (P, 2),
              "rotation": (P, 4)}
    groups = [{"params": [torch.nn.Parameter(torch.randn(s, generator=g))], "lr": 0.01, "name": k}
              for k, s in shapes.items()]
    groups.append({"params": [torch.nn.Parameter(torch.randn(7, generator=g))], "lr": 0.01, "name": "other"})
    opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for _ in range(2):
        for grp in opt.param_groups:
            grp["params"][0].grad = torch.randn(grp["params"][0].shape, generator=g)
        opt.step()
    return opt


def test_rewrap_moves_state_like_the_reference():
    opt = _optimizer()
    old = {g["name"]: g["params"][0] for g in opt.param_groups}
    keep = torch.arange(50) % 3 != 0
    new_params = {k: old[k].detach()[keep].clone() for k in densify.PARAM_NAMES}
    new_moments = {k: (opt.state[old[k]]["exp_avg"][keep].clone(), opt.state[old[k]]["exp_avg_sq"][keep].clone())
                   for k in densify.PARAM_NAMES}
    steps = {k: opt.state[old[k]]["step"] for k in densify.PARAM_NAMES}
    out = densify.rewrap_optimizer(opt, new_params, new_moments)
    assert set(out) == set(densify.PARAM_NAMES)
    assert len(opt.state) == 7  # old keys gone, new keys in, the unrelated group untouched
    for g in opt.param_groups:
        k, p = g["name"], g["params"][0]
        if k == "other":
            assert p is old[k]
            continue
        assert p is out[k] and p is not old[k] and isinstance(p, torch.nn.Parameter) and p.requires_grad
        assert old[k] not in opt.state
        st = opt.state[p]
        assert st["step"] is steps[k]
        assert st["exp_avg"] is new_moments[k][0] and st["exp_avg_sq"] is new_moments[k][1]
        p.grad = torch.ones_like(p)
    opt.step()  # shapes are consistent: Adam keeps running on the compacted tensors


def test_rewrap_without_state_and_error_paths():
    P = 5
    prm = {k: torch.nn.Parameter(torch.zeros((P,) + s)) for k, s in
           {"xyz": (3,), "f_dc": (1, 3), "f_rest": (0, 3), "opacity": (1,), "scaling": (2,), "rotation": (4,)}.items()}
    opt = torch.optim.Adam([{"params": [v], "name": k} for k, v in prm.items()], lr=0.0)
    out = densify.rewrap_optimizer(opt, {k: torch.ones((2,) + tuple(v.shape[1:])) for k, v in prm.items()},
                                   {k: None for k in prm})
    assert all(out[k].shape[0] == 2 for k in prm) and len(opt.state) == 0
    opt2 = _optimizer()
    with pytest.raises(RuntimeError, match="no new moments"):
        densify.rewrap_optimizer(opt2, {"xyz": torch.zeros(4, 3)}, {})


def test_cpu_tensors_are_rejected_not_silently_processed(lib):
    params = {k: torch.zeros((4,) + s) for k, s in
              {"xyz": (3,), "f_dc": (1, 3), "f_rest": (3, 3), "opacity": (1,), "scaling": (2,), "rotation": (4,)}.items()}
    with pytest.raises(RuntimeError, match="CUDA"):
        densify.densify_and_prune(params, {k: None for k in params}, None, torch.zeros(4, 1), torch.zeros(4, 1),
                                  0.0002, 0.005, 1.0, 20, 0.01)


def test_names_agree_with_the_oracle():
    assert densify.PARAM_NAMES == densify_oracle.PARAM_NAMES


def test_abi_argument_validation(lib):
    assert lib.pgs_densify_blocks(0) == 0 and lib.pgs_densify_blocks(257) == 2
    assert lib.pgs_densify_plan(-1, None, None, None, None, 0.0, 0.0, 0.0, 0, 0.0, 1.6, None, None, None, None) < 0
    assert lib.pgs_densify_plan(4, None, None, None, None, 0.0, 0.0, 0.0, 0, 0.0, 1.6, None, None, None, None) < 0
    assert lib.pgs_densify_gather(25, None, None, None, None, 1, 1, None, None) < 0
    assert b"at most 24" in lib.pgs_last_error()
    assert lib.pgs_densify_gather(0, None, None, None, None, 0, 0, None, None) == 0
    assert lib.pgs_densify_children(0, None, None, None, None, None, None, None, 1.6, None, None, None) == 0
    assert lib.pgs_densify_map(3, None, None, None, 0, None, None, None) < 0
