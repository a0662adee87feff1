"""Parity at the sizes BASELINE.json names (configs 1-5): full-size runs checked against the C oracle where it
finishes in seconds, and through size-independent properties (orthonormal rotations, linearity of the reverse
mode in the upstream gradient, shard invariance) where it does not."""
import numpy as np
import pytest
import torch

from conftest import assert_differences_are_near_ties, get_desc, load_golden, planning_cost_from_golden, rel_err, stretch_cost_from_golden
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


def uniform_q(d, n, seed):
    rng = np.random.default_rng(seed)
    lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
    hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
    return (lo + rng.random((n, d.n_dofs), dtype=np.float32) * (hi - lo)).astype(np.float32)


def test_config1_panda_fk_1024_rand01():
    """Config 1: torch.rand(1024, 7) as in examples/forward_kinematics.py (joint 4 always clamped)."""
    from torch_robotics_b200.torch_kinematics_tree.models.robots import DifferentiableFrankaPanda
    d = get_desc("panda")
    q = torch.rand(1024, 7, generator=torch.Generator().manual_seed(1))
    panda = DifferentiableFrankaPanda(device=DEV)
    qd = q.to(DEV).requires_grad_(True)
    H = panda.compute_forward_kinematics_all_links(qd)
    Ho = co.fk_forward(co.Tree(d), q.numpy())
    assert rel_err(H.detach().cpu().numpy(), Ho) < 1e-5
    G = torch.randn(H.shape, generator=torch.Generator().manual_seed(2))
    (H * G.to(DEV)).sum().backward()
    go = co.fk_backward(co.Tree(d), q.numpy(), G.numpy())
    assert rel_err(qd.grad.cpu().numpy(), go) < 1e-4
    assert torch.all(qd.grad[:, 3] == 0)            # clamped joint: exactly zero gradient (torch.clamp semantics)


def test_config2_panda_fk_backward_1M():
    """Config 2: 1,048,576 configs.  Full-size check against the C oracle (OpenMP) + properties."""
    from torch_robotics_b200._lib import DeviceModel
    d = get_desc("panda")
    m = DeviceModel(d, DEV)
    n = 1 << 20
    q = uniform_q(d, n, 7)
    qd = t(q)
    H = m.fk_forward(qd)
    Ho = co.fk_forward(co.Tree(d), q)
    assert rel_err(H.cpu().numpy(), Ho) < 1e-5
    R = H[:, :, :3, :3]
    eye = torch.eye(3, device=DEV)
    assert float((R @ R.transpose(-1, -2) - eye).abs().max()) < 2e-6          # rotations stay orthonormal
    assert float((torch.linalg.det(R[::97]) - 1).abs().max()) < 2e-6
    gen = torch.Generator(device=DEV).manual_seed(3)
    G1 = torch.randn(H.shape, device=DEV, generator=gen)
    G2 = torch.randn(H.shape, device=DEV, generator=gen)
    g1, g2 = m.fk_backward(qd, G1), m.fk_backward(qd, G2)
    g12 = m.fk_backward(qd, 2.0 * G1 - 0.5 * G2)
    ref = 2.0 * g1 - 0.5 * g2
    assert float((g12 - ref).abs().max() / ref.abs().max()) < 1e-5            # linear in the upstream gradient
    sub = slice(0, 65536)
    go = co.fk_backward(co.Tree(d), q[sub], G1[sub].cpu().numpy())
    assert rel_err(g1[sub].cpu().numpy(), go) < 1e-4
    # shard invariance: any contiguous shard gives bit-identical rows (what multi-GPU sharding relies on)
    lo, hi = 300_001, 777_777
    assert torch.equal(m.fk_forward(qd[lo:hi].contiguous()), H[lo:hi])
    assert torch.equal(m.fk_backward(qd[lo:hi].contiguous(), G1[lo:hi].contiguous()), g1[lo:hi])


def test_config3_planning_objective_4096x64():
    """Config 3: 4,096 trajectories x 64 waypoints, self + object + ws + EE, cost and gradient vs the C oracle."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g, d = load_golden("planning_panda_spheres3d"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    m = DeviceModel(d, DEV)
    dc = DeviceCost(m, cost)
    q = uniform_q(d, 4096 * 64, 11)
    r = dc.objective(t(q), want_grad=True, want_terms=True, want_occupancy=True, want_argmin=True)
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    assert rel_err(r["terms"].cpu().numpy(), o["terms"]) < 1e-5
    assert rel_err(r["total"].cpu().numpy(), o["total"]) < 1e-5
    # gradients: a configuration sitting on a decision boundary (two spheres / half-spaces equally near within
    # an ulp of the link positions) legitimately routes its gradient to the other candidate on one side;
    # such rows are counted, everything else must agree to 1e-4
    gd = np.abs(r["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    flipped = gd >= 1e-4
    print(f"config 3: gradient rows on a decision boundary: {int(flipped.sum())} of {len(gd)}")
    assert gd[~flipped].max() < 1e-4
    # every difference is proven to be a near-tie: within a few ulp of a decision boundary on the oracle's own positions
    n_diff = assert_differences_are_near_ties(d, cost, q, r["argmin"].cpu().numpy(), o["argmin"], r["occupancy"].cpu().numpy(), o["occ"],
                                              np.flatnonzero(flipped), "config 3")
    assert n_diff <= 30
    # the (b, h) shaped public entry point gives the same numbers
    from torch_robotics_b200 import ops
    c = ops.objective(dc, t(q))
    assert torch.equal(c, r["total"])


def test_config5_stretch_objective_2M():
    """Config 5 at 2,097,152 configs on one GPU (16M = 8 such shards): vs the C oracle on a 131,072 subset,
    shard invariance on the rest."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g, d = load_golden("objective_stretch"), get_desc("stretch")
    cost = stretch_cost_from_golden(g)
    m = DeviceModel(d, DEV)
    dc = DeviceCost(m, cost)
    n = 1 << 21
    q = uniform_q(d, n, 5)
    r = dc.objective(t(q), want_grad=True, want_terms=True)
    sub = slice(1_000_000, 1_000_000 + 131072)
    o = co.objective(co.Tree(d), co.Cost(cost), q[sub])
    assert rel_err(r["total"][sub].cpu().numpy(), o["total"]) < 1e-5
    gd = np.abs(r["grad_q"][sub].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert gd[gd < 1e-4].max() < 1e-4
    n_diff = assert_differences_are_near_ties(d, cost, q[sub], None, None, None, None, np.flatnonzero(gd >= 1e-4), "config 5")
    assert n_diff <= 30
    r2 = dc.objective(t(q[sub]), want_grad=True, want_terms=False)
    assert torch.equal(r2["total"], r["total"][sub]) and torch.equal(r2["grad_q"], r["grad_q"][sub])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_bit_identical_to_one():
    """Row-sharding over devices does not change any per-configuration result."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.distributed import shard_bounds
    g, d = load_golden("planning_panda_spheres3d"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    q = uniform_q(d, 200_003, 13)
    full = DeviceCost(DeviceModel(d, "cuda:0"), cost).objective(torch.as_tensor(q, device="cuda:0"))
    for rank in range(2):
        lo, hi = shard_bounds(q.shape[0], rank, 2)
        dev = f"cuda:{rank}"
        part = DeviceCost(DeviceModel(d, dev), cost).objective(torch.as_tensor(q[lo:hi], device=dev))
        assert torch.equal(part["total"].cpu(), full["total"][lo:hi].cpu())
        assert torch.equal(part["grad_q"].cpu(), full["grad_q"][lo:hi].cpu())


def test_fused_cost_sum_matches_sum_of_costs():
    """trk_objective_sums: the per-block partial sums the kernel writes add up to sum(cost_total) (fp32 partials: 1e-6
    relative), the other outputs are bit-identical to trk_objective's, the result is reproducible, and a model that is not
    served by the headline kernel reports TRK_E_UNSUPPORTED."""
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    g, d = load_golden("planning_panda_spheres3d"), get_desc("panda")
    cost = planning_cost_from_golden(g)
    dc = DeviceCost(DeviceModel(d, DEV), cost)
    for n in (1, 777, 300_001):
        q = t(uniform_q(d, n, 3))
        ref = dc.objective(q, want_grad=True, want_terms=False)
        grad, tot = torch.empty_like(q), torch.empty(n, device=DEV)
        bs = dc.objective_sum(q, out_grad=grad, out_total=tot)
        assert torch.equal(tot, ref["total"]) and torch.equal(grad, ref["grad_q"])
        s, s_ref = float(bs.sum(dtype=torch.float64)), float(ref["total"].sum(dtype=torch.float64))
        assert abs(s - s_ref) <= 1e-5 * max(1.0, float(ref["total"].abs().sum(dtype=torch.float64)))
        bs2 = dc.objective_sum(q, out_grad=grad)                       # cost_total is optional; same partials again
        assert torch.equal(bs, bs2)
    dk = get_desc("kuka_iiwa")
    from torch_robotics_b200.descriptor import build_cost_descriptor
    ck = build_cost_descriptor(pt_link=[3, 5], pt_radius=[0.1, 0.1], cutoff_obj=0.01,
                               objects=[(np.array([[0.4, 0.0, 0.5]], np.float32), np.array([0.1], np.float32), None, None)],
                               ws_min=[-1, -1, -1], ws_max=[1, 1, 1])
    dck = DeviceCost(DeviceModel(dk, DEV), ck)
    qk = t(uniform_q(dk, 64, 1))
    with pytest.raises(RuntimeError, match="fused cost reduction"):
        dck.objective_sum(qk, out_grad=torch.empty_like(qk))
