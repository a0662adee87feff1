"""analytic link Jacobians (DifferentiableTree.compute_analytical_jacobian_all_links,
robot_tree.py:250-265).  Golden: tests/golden/link_jacobians.npz, the reference's autograd Jacobian."""
import numpy as np
import pytest
import torch

from conftest import get_desc, load_golden, rel_err
from oracle import restatement as orc

ROBOTS = ["panda", "stretch", "shadow_hand"]


def quat_sign_aligned(J, Jref, Rw):
    """rotation_matrix_to_q picks the candidate with the largest component; where two components tie structurally
    (e.g. a 90-degree joint origin makes |w| == |x| for every q) rounding noise decides and the whole quaternion -
    hence rows 3-6 - may come out negated.  Align the sign of those blocks, and only those."""
    R = np.asarray(Rw, np.float64)
    t = np.stack([1 + R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2], 1 + R[..., 0, 0] - R[..., 1, 1] - R[..., 2, 2],
                  1 - R[..., 0, 0] + R[..., 1, 1] - R[..., 2, 2], 1 - R[..., 0, 0] - R[..., 1, 1] + R[..., 2, 2]], -1)
    qa = np.sort(np.sqrt(np.maximum(t, 0)), -1)
    tie = (qa[..., 3] - qa[..., 2]) < 1e-4                                   # [N, L]
    J = J.copy()
    flip = tie & ((J[:, :, 3:7] * Jref[:, :, 3:7]).sum((-1, -2)) < 0)
    J[:, :, 3:7] = np.where(flip[:, :, None, None], -J[:, :, 3:7], J[:, :, 3:7])
    return J, int(flip.sum())


@pytest.mark.parametrize("robot", ROBOTS)
def test_oracle_matches_reference(robot):
    g, d = load_golden("link_jacobians"), get_desc(robot)
    q, Jref = g[f"{robot}_q"], g[f"{robot}_J"]
    J = orc.link_jacobian(d, q, np.float32)
    assert J.shape == Jref.shape == (q.shape[0], d.n_links, 7, d.n_dofs)
    J, _ = quat_sign_aligned(J, Jref, g[f"{robot}_H"][..., :3, :3])
    assert rel_err(J, Jref) < 1e-5
    # clamped joints (out-of-limit rows) have exactly zero columns, like torch.clamp's backward
    _, mask = orc.joint_angles(d, q, np.float32)
    for b in range(d.n_links):
        if d.q_index[b] >= 0:
            assert np.all(J[mask[:, b] == 0][:, :, :, d.q_index[b]] == 0)


def test_jacobian_matches_finite_differences():
    d = get_desc("stretch")
    rng = np.random.default_rng(0)
    q = rng.uniform(-0.3, 0.3, (4, d.n_dofs))
    q[:, [i for i, b in enumerate(d.controlled) if d.has_lim[b]]] = 0.05      # inside every limit
    J = orc.link_jacobian(d, q, np.float64)
    eps = 1e-6
    for j in range(d.n_dofs):
        dq = np.zeros_like(q); dq[:, j] = eps
        Rp, pp = orc.fk_forward(d, q + dq, np.float64)
        Rm, pm = orc.fk_forward(d, q - dq, np.float64)
        assert np.abs((pp - pm) / (2 * eps) - J[:, :, 0:3, j]).max() < 1e-6
        qp, qm = orc.rotation_to_quaternion(Rp, np.float64), orc.rotation_to_quaternion(Rm, np.float64)
        ok = (qp * qm).sum(-1) > 0.99                                             # same representative on both sides
        assert np.abs(((qp - qm) / (2 * eps) - J[:, :, 3:7, j])[ok]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ROBOTS)
def test_gpu_matches_reference_and_oracle(robot):
    from torch_robotics_b200.torch_kinematics_tree.models import robots as R
    cls = dict(panda=lambda: R.DifferentiableFrankaPanda(gripper=False, device="cuda:0"),
               stretch=lambda: R.DifferentiableHabitatStretch(device="cuda:0"),
               shadow_hand=lambda: R.DifferentiableShadowHand(device="cuda:0"))[robot]
    g, d = load_golden("link_jacobians"), get_desc(robot)
    model = cls()
    q = torch.tensor(g[f"{robot}_q"], device="cuda:0")
    J = model.compute_analytical_jacobian_all_links(q).cpu().numpy()
    Ja, _ = quat_sign_aligned(J, g[f"{robot}_J"], g[f"{robot}_H"][..., :3, :3])
    assert rel_err(Ja, g[f"{robot}_J"]) < 1e-5
    # larger batch (ragged tail: not a multiple of the 32-configuration warp tile) against the numpy restatement
    rng = np.random.default_rng(11)
    lo = np.where(d.has_lim[d.controlled] > 0, d.lo[d.controlled], -np.pi).astype(np.float32)
    hi = np.where(d.has_lim[d.controlled] > 0, d.hi[d.controlled], np.pi).astype(np.float32)
    qn = (lo + rng.random((1000 + 13, d.n_dofs), dtype=np.float32) * (hi - lo) * 1.1 - 0.05 * (hi - lo)).astype(np.float32)
    Jg = model.device_model("cuda:0").link_jacobian(torch.as_tensor(qn, device="cuda:0")).cpu().numpy()
    Jo = orc.link_jacobian(d, qn, np.float32)
    Jg, flips = quat_sign_aligned(Jg, Jo, orc.fk_forward(d, qn, np.float32)[0])
    assert rel_err(Jg, Jo) < 1e-5
    assert model.compute_analytical_jacobian_all_links(q[0]).shape == (1, d.n_links, 7, d.n_dofs)


@pytest.mark.gpu
def test_gpu_link_list_and_attached_points():
    """link_list selects and orders the links; attached points get Jacobians like any other body."""
    from torch_robotics_b200.torch_kinematics_tree.models import robots as R
    model = R.DifferentiableFrankaPanda(gripper=False, device="cuda:0", link_list=["ee_link", "panda_link3"])
    q = torch.rand(5, 7, device="cuda:0") * 0.5 - 1.0
    J = model.compute_analytical_jacobian_all_links(q)
    full = R.DifferentiableFrankaPanda(gripper=False, device="cuda:0")
    Jf = full.compute_analytical_jacobian_all_links(q)
    assert torch.equal(J, Jf[:, [10, 3]])
    _, ids = full.attach_points([("panda_hand", (0.05, 0.0, 0.1))])
    Jp = full.device_model("cuda:0").link_jacobian(q).cpu().numpy()
    Jo = orc.link_jacobian(full._desc, q.cpu().numpy(), np.float32)
    assert rel_err(Jp[:, ids[0], :3], Jo[:, ids[0], :3]) < 1e-5


@pytest.mark.gpu
def test_geometric_jacobian_matches_reference():
    """compute_forward_kinematics_and_geometric_jacobian (robot_tree.py:218-248) against the reference's own output:
    position, wxyz quaternion, linear and angular Jacobian of a link - Panda (chain), UR10 (y / z axes), dual-arm Tiago
    (branches: the reference's joint-index rule admits joints that are not ancestors of the link)."""
    import torch
    from torch_robotics_b200.torch_kinematics_tree.models import robots as R
    g = load_golden("geometric_jacobian")
    for name, cls in (("panda", R.DifferentiableFrankaPanda), ("ur10", R.DifferentiableUR10), ("tiago_dual_holo", R.DifferentiableTiagoDualHolo)):
        tree = cls(device="cuda:0")
        q = torch.as_tensor(g[f"{name}_q"], device="cuda:0")
        links = [str(x) for x in g[f"{name}_links"]]
        for ln in links:
            pos, rot, lin, ang = tree.compute_forward_kinematics_and_geometric_jacobian(q, torch.zeros_like(q), ln)
            assert lin.shape == (q.shape[0], 3, tree._n_dofs) and rot.shape == (q.shape[0], 4)
            assert rel_err(pos.cpu().numpy(), g[f"{name}_{ln}_pos"]) < 1e-5, (name, ln)
            # a quaternion and its negative are the same rotation; the branch choice can flip on a near-tie of the trace test
            rq, rg = rot.cpu().numpy(), g[f"{name}_{ln}_rot"]
            sgn = np.sign(np.sum(rq * rg, axis=1, keepdims=True))
            assert rel_err(rq * sgn, rg) < 1e-5, (name, ln)
            assert rel_err(lin.cpu().numpy(), g[f"{name}_{ln}_lin"]) < 1e-5, (name, ln)
            assert rel_err(ang.cpu().numpy(), g[f"{name}_{ln}_ang"]) < 1e-5, (name, ln)
            zc_ref = np.all(g[f"{name}_{ln}_ang"] == 0, axis=(0, 1))
            zc = np.all(ang.cpu().numpy() == 0, axis=(0, 1))
            assert np.array_equal(zc, zc_ref), f"{name}/{ln}: excluded q columns differ"
        p2, r2 = tree.compute_forward_kinematics(q, None, links[0], state_less=False)
        assert rel_err(p2.cpu().numpy(), g[f"{name}_{links[0]}_pos"]) < 1e-5
