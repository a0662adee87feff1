"""torch.autograd.Function wrappers over the C ABI (the drop-in seam of BASELINE.json's north star).

  FKFunction         q -> H [N, n_sel, 4, 4]          forward: trk_fk_forward, backward: trk_fk_backward
  ObjectiveFunction  q -> total cost [N]              one fused launch computes the cost AND d cost / d q
  CostStageFunction  link_pos (, ee_pose) -> cost [N] the DistanceField.compute_cost seam (FK output as a tensor)

Reference counterparts: autograd over models/rigid_body.py:146-211 and over
torch_planning_objectives/fields/distance_fields.py:107-356.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ._lib import DeviceCost, DeviceModel


class FKFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, model: DeviceModel, link_sel):
        qc = q.detach()
        H = model.fk_forward(qc, link_sel)
        ctx.model, ctx.link_sel = model, link_sel
        ctx.save_for_backward(qc)
        return H

    @staticmethod
    def backward(ctx, grad_H):
        (q,) = ctx.saved_tensors
        return ctx.model.fk_backward(q, grad_H.contiguous(), ctx.link_sel), None, None


def fk(model: DeviceModel, q: torch.Tensor, link_sel: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Homogeneous poses [N, n_sel, 4, 4] of the selected bodies; differentiable w.r.t. q."""
    if q.requires_grad and torch.is_grad_enabled():
        return FKFunction.apply(q, model, None if link_sel is None else tuple(int(i) for i in link_sel))
    return model.fk_forward(q.detach(), link_sel)


class ObjectiveFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, cost: DeviceCost, ee_target_cfg):
        r = cost.objective(q.detach(), want_grad=True, want_terms=False, ee_target_cfg=ee_target_cfg)
        ctx.save_for_backward(r["grad_q"])
        return r["total"]

    @staticmethod
    def backward(ctx, g):
        (grad_q,) = ctx.saved_tensors
        return grad_q * g.reshape(-1, 1), None, None


def objective(cost: DeviceCost, q: torch.Tensor, ee_target_cfg: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weighted total cost [N]; the gradient w.r.t. q comes out of the same kernel launch."""
    if q.requires_grad and torch.is_grad_enabled():
        return ObjectiveFunction.apply(q, cost, ee_target_cfg)
    return cost.objective(q.detach(), want_grad=False, want_terms=False, ee_target_cfg=ee_target_cfg)["total"]


class CostStageFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, link_pos, ee_pose, cost: DeviceCost):
        r = cost.cost_stage(link_pos.detach(), None if ee_pose is None else ee_pose.detach(), want_grad=True)
        ctx.has_ee = ee_pose is not None
        ctx.save_for_backward(r["grad_pos"], r["grad_ee"] if ctx.has_ee else r["grad_pos"].new_empty(0))
        return r["total"]

    @staticmethod
    def backward(ctx, g):
        gpos, gee = ctx.saved_tensors
        g = g.reshape(-1)
        return gpos * g[:, None, None], (gee * g[:, None, None] if ctx.has_ee else None), None


def cost_stage(cost: DeviceCost, link_pos: torch.Tensor, ee_pose: Optional[torch.Tensor] = None) -> torch.Tensor:
    needs = torch.is_grad_enabled() and (link_pos.requires_grad or (ee_pose is not None and ee_pose.requires_grad))
    if needs:
        return CostStageFunction.apply(link_pos, ee_pose, cost)
    return cost.cost_stage(link_pos.detach(), None if ee_pose is None else ee_pose.detach(), want_grad=False)["total"]
