"""Fused optimiser step and densification bookkeeping (the per-iteration part).

``FusedAdam`` is a drop-in for the ``torch.optim.Adam(l, lr=0.0, eps=1e-15)`` PartGS builds in
``GaussianModel.training_setup`` (scene/gaussian_model.py:256-266): same param-group interface (``lr`` per group is
what the reference's schedulers write), same ``state[p] = {step, exp_avg, exp_avg_sq}`` layout (so the reference's
``_prune_optimizer`` / ``cat_tensors_to_optimizer`` keep working), same arithmetic — but one kernel launch per step
for all groups (csrc/optim.cu, ``pgs_adam_step``).  ``densification_stats`` fuses ``add_densification_stats`` and the
``max_radii2D`` update.  CUDA float32 tensors only; anything else raises.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib

_MAX = 16


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        # tensors that share (step, betas, eps) go into one launch
        batches = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (_lib.on_device(p) and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                    raise RuntimeError("FusedAdam: CUDA float32 parameters and gradients only (no fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                key = (p.device, st["step"], beta1, beta2, group["eps"])
                batches.setdefault(key, []).append((p, p.grad.contiguous(), st, group["lr"]))
        for (dev, step, beta1, beta2, eps), items in batches.items():
            bias_correction1 = 1 - beta1 ** step
            bias_correction2_sqrt = (1 - beta2 ** step) ** 0.5
            for s0 in range(0, len(items), _MAX):
                chunk = items[s0:s0 + _MAX]
                n = len(chunk)
                for p, g, st, _ in chunk:
                    if not (p.is_contiguous() and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()):
                        raise RuntimeError("FusedAdam: parameters and optimiser state must be contiguous")
                pa = (C.c_void_p * n)(*[p.data_ptr() for p, _, _, _ in chunk])
                ga = (C.c_void_p * n)(*[g.data_ptr() for _, g, _, _ in chunk])
                ma = (C.c_void_p * n)(*[st["exp_avg"].data_ptr() for _, _, st, _ in chunk])
                va = (C.c_void_p * n)(*[st["exp_avg_sq"].data_ptr() for _, _, st, _ in chunk])
                na = (C.c_size_t * n)(*[p.numel() for p, _, _, _ in chunk])
                sa = (C.c_float * n)(*[lr / bias_correction1 for _, _, _, lr in chunk])
                with torch.cuda.device(dev):
                    rc = lib.pgs_adam_step(n, pa, ga, ma, va, na, sa, float(beta1), float(beta2), float(eps),
                                           float(bias_correction2_sqrt), _lib.current_stream(dev))
                _lib.check(rc, "pgs_adam_step")
        return loss


def densification_stats(radii, viewspace_grad, xyz_gradient_accum, denom, max_radii2D=None):
    """In place, for ``visibility_filter = radii > 0``: ``max_radii2D = max(max_radii2D, radii)``,
    ``xyz_gradient_accum += ||viewspace_grad[:, :2]||``, ``denom += 1`` (train.py:295-297 +
    scene/gaussian_model.py:515-517)."""
    lib = _lib.load()
    P = radii.numel()
    if not (_lib.on_device(radii) and radii.dtype == torch.int32):
        raise RuntimeError("radii must be a CUDA int32 tensor")
    g = _lib.require_cuda_float(viewspace_grad, "viewspace_grad")
    for t, nm in ((xyz_gradient_accum, "xyz_gradient_accum"), (denom, "denom"), (max_radii2D, "max_radii2D")):
        if t is not None and not (_lib.on_device(t) and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == P):
            raise RuntimeError(f"{nm} must be a contiguous CUDA float32 tensor with one entry per surfel")
    if g.dim() != 2 or g.size(0) != P or g.size(1) < 3 or g.stride(0) != 3:
        raise RuntimeError("viewspace_grad must be [P,3] contiguous")
    with torch.cuda.device(radii.device):
        rc = lib.pgs_densify_stats(P, radii.data_ptr(), g.data_ptr(), _lib.ptr(max_radii2D),
                                   xyz_gradient_accum.data_ptr(), denom.data_ptr(), _lib.current_stream(radii.device))
    _lib.check(rc, "pgs_densify_stats")


class ShardedFusedAdam:
    """Data-parallel Adam with the optimiser state sharded over the ranks (SURVEY.md §8(e) "alternative to evaluate",
    §8(f) rank 3): instead of all-reducing the 232 B/surfel gradient bucket and running the same Adam step on every
    rank, the gradients are **reduce-scattered**, every rank updates only its 1/W slice of the parameters (and keeps
    only that slice of ``exp_avg / exp_avg_sq``: 1/W of the 464 B/surfel optimiser state), and the updated slices are
    **all-gathered**.  Same bytes on the wire as an all-reduce, 1/W of the optimiser work and state per rank.

    All parameters live in ONE flat float32 buffer (``self.flat``); the tensors handed in are re-pointed at views of
    it (``p.data``), and ``p.grad`` at views of a second flat buffer that autograd accumulates into in place — so
    the collectives move one contiguous range each and no gather / scatter copies exist.  The update itself is the
    library's one-launch Adam (``pgs_adam_step``): one table entry per (parameter ∩ my slice), each with its group's
    learning rate.  ``param_groups`` keeps the ``{"name", "lr"}`` entries the reference's schedulers write to
    (scene/gaussian_model.py:268-275); it is NOT interchangeable with torch.optim.Adam beyond that (fixed-size models;
    see `state`, `state_dict`, `rebuild` below).  Arithmetic = torch.optim.Adam(eps=1e-15) on the rank-summed gradient.
    """

    def __init__(self, named_params, betas=(0.9, 0.999), eps=1e-15, group=None, average: bool = False):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.betas, self.eps, self.average = betas, eps, average
        self.param_groups = []
        params = []
        for g in named_params:
            if len(g["params"]) != 1:
                raise RuntimeError("ShardedFusedAdam: one tensor per group (as in training_setup)")
            p = g["params"][0]
            if not (_lib.on_device(p) and p.dtype == torch.float32):
                raise RuntimeError("ShardedFusedAdam: CUDA float32 parameters only (no fallback)")
            params.append(p)
            self.param_groups.append({"name": g.get("name"), "lr": float(g.get("lr", 0.0)), "params": [p]})
        dev = params[0].device
        align = 64 * self.world
        self.offsets, total = [], 0
        for p in params:
            self.offsets.append(total)
            total += (p.numel() + 63) // 64 * 64
        self.numel = (total + align - 1) // align * align
        self.slice = self.numel // self.world
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        lo = self.rank * self.slice
        self.exp_avg = torch.zeros(self.slice, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.slice, dtype=torch.float32, device=dev)
        self.grad_slice = torch.zeros(self.slice, dtype=torch.float32, device=dev)
        self.step_count = 0
        self._lo = lo

    # ---- what this optimiser is NOT: a drop-in for model surgery --------------------------------------------------
    # It owns flat parameter / gradient buffers and re-points `p.data` / `p.grad` at views of them, and it keeps only
    # 1/W of the moments.  Densification (densify.rewrap_optimizer, prune / cat of the reference) replaces the
    # Parameters and resizes per-parameter state: rebuild a new ShardedFusedAdam from the new tensors afterwards
    # (`rebuild`), the moments of surviving rows are carried over by the caller through state_dict / load_state_dict
    # of the gathered state.  `state` (torch's per-parameter dict) does not exist here, on purpose.
    @property
    def state(self):
        raise RuntimeError("ShardedFusedAdam keeps ONE sharded moment buffer, not torch's per-parameter `state`: use "
                           "state_dict() / load_state_dict(), and rebuild() after the parameter tensors were replaced "
                           "(densification); for optimiser-state surgery use FusedAdam")

    def state_dict(self):
        """This rank's shard: {'step', 'rank', 'world', 'numel', 'exp_avg', 'exp_avg_sq', 'param_groups': [{name, lr}]}.
        A checkpoint of the whole optimiser is the list of the W shards (torch.distributed.all_gather_object, or
        one file per rank)."""
        return {"step": self.step_count, "rank": self.rank, "world": self.world, "numel": self.numel,
                "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{"name": g["name"], "lr": g["lr"]} for g in self.param_groups]}

    def load_state_dict(self, sd):
        if sd["world"] != self.world or sd["rank"] != self.rank or sd["numel"] != self.numel:
            raise RuntimeError("ShardedFusedAdam.load_state_dict: shard of another layout "
                               f"(rank {sd['rank']}/{sd['world']}, {sd['numel']} elements)")
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, h in zip(self.param_groups, sd["param_groups"]):
            g["lr"] = float(h["lr"])

    def rebuild(self, named_params):
        """A fresh optimiser over new parameter tensors (after densification replaced them), same hyper-parameters;
        moments start at zero like the reference's cat_tensors_to_optimizer does for new rows."""
        return ShardedFusedAdam(named_params, betas=self.betas, eps=self.eps, group=self.group, average=self.average)

    def zero_grad(self, set_to_none: bool = False):
        """Gradients are views of the flat buffer that autograd accumulates into in place: they are zeroed and
        re-attached, never dropped (``set_to_none`` is accepted for signature compatibility and ignored)."""
        self.flat_grad.zero_()
        for g, o in zip(self.param_groups, self.offsets):
            p = g["params"][0]
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    @torch.no_grad()
    def step(self):
        lib = _lib.load()
        dist, W = self._dist, self.world
        lo, hi = self._lo, self._lo + self.slice
        # 1. my slice of the rank-summed gradient
        if W > 1:
            backend = dist.get_backend(self.group)
            if backend == "nccl":
                dist.reduce_scatter_tensor(self.grad_slice, self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            else:  # gloo (CPU tests) has no reduce-scatter: all-reduce, then take the slice
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
                self.grad_slice.copy_(self.flat_grad[lo:hi])
            if self.average:
                self.grad_slice.div_(W)
        else:
            self.grad_slice.copy_(self.flat_grad[lo:hi])
        # 2. one-launch Adam over (parameter ∩ my slice), each piece with its group's learning rate
        self.step_count += 1
        beta1, beta2 = self.betas
        bc1 = 1 - beta1 ** self.step_count
        bc2_sqrt = (1 - beta2 ** self.step_count) ** 0.5
        pieces = []
        for g, o in zip(self.param_groups, self.offsets):
            a, b = max(o, lo), min(o + g["params"][0].numel(), hi)
            if a < b:
                pieces.append((a, b - a, g["lr"] / bc1))
        dev = self.flat.device
        base_p, base_g = self.flat.data_ptr(), self.grad_slice.data_ptr()
        base_m, base_v = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        for s0 in range(0, len(pieces), _MAX):
            chunk = pieces[s0:s0 + _MAX]
            n = len(chunk)
            pa = (C.c_void_p * n)(*[base_p + 4 * a for a, _, _ in chunk])
            ga = (C.c_void_p * n)(*[base_g + 4 * (a - lo) for a, _, _ in chunk])
            ma = (C.c_void_p * n)(*[base_m + 4 * (a - lo) for a, _, _ in chunk])
            va = (C.c_void_p * n)(*[base_v + 4 * (a - lo) for a, _, _ in chunk])
            na = (C.c_size_t * n)(*[cnt for _, cnt, _ in chunk])
            sa = (C.c_float * n)(*[ss for _, _, ss in chunk])
            with torch.cuda.device(dev):
                rc = lib.pgs_adam_step(n, pa, ga, ma, va, na, sa, float(beta1), float(beta2), float(self.eps),
                                       float(bc2_sqrt), _lib.current_stream(dev))
            _lib.check(rc, "pgs_adam_step")
        # 3. everybody gets every updated slice
        if W > 1:
            dist.all_gather_into_tensor(self.flat, self.flat[lo:hi].clone(), group=self.group)
