"""Run-time specialisation (torch_robotics_b200/specialize.py): a robot outside the shipped set gets a private build of
the library with its tree compiled in.  CPU: the generated header, the build (nvcc cross-compiles without a GPU), the
exports of the private library, error behaviour.  GPU: the specialised kernels against the oracle and against the
runtime-tree kernels of the default library on the same inputs."""
import os
import re

import numpy as np
import pytest

from conftest import get_desc, rel_err
import specialized_cases as sc

POSE_TOL, GRAD_TOL = 1e-5, 1e-4


def test_header_describes_the_descriptor():
    from torch_robotics_b200 import specialize as sp
    d = sc.iiwa7_allegro_desc()
    h = sp.specialized_header(d)
    assert re.search(r"struct Jit_[0-9a-f]{12} \{", h) and "using Jit = Jit_" in h and f"static constexpr int L = {d.n_links}, D = {d.n_dofs};" in h
    assert "#define TRK_FOR_EACH_STATIC_ROBOT(X) X(robots::Jit)" in h and "TRK_JIT_HAS_COLLISION" not in h
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{8})u", h)]
    assert words == [w & 0xFFFFFFFF for w in sp.tree_words(d)]          # the image trk_model_create compares a model with
    body = [int(v) for v in re.search(r"int body\[L\] = \{([^}]*)\}", h).group(1).split(",")]
    assert body == [int(b) for b in d.topo]
    # the hexadecimal literals are the descriptor's floats, bit for bit
    lo = [float.fromhex(v.rstrip("f")) for v in re.search(r"float lo\[L\] = \{([^}]*)\}", h).group(1).split(", ")]
    assert np.array_equal(np.asarray(lo, np.float32), d.lo[d.topo])
    d2, cm = sc.ur10_collision()
    h2 = sp.specialized_header(d2, cm)
    assert re.search(r"struct JitCollision_[0-9a-f]{12} \{", h2) and "#define TRK_JIT_HAS_COLLISION 1" in h2
    assert f"static constexpr int S = {len(cm['pt_link'])}, P = {len(cm['pairs'])};" in h2
    # the library's name follows the content: another robot, or the same robot with a collision model, is another file
    assert len({sp.library_path(d), sp.library_path(d2), sp.library_path(d2, cm)}) == 3


def test_specialize_errors():
    import copy
    from torch_robotics_b200 import specialize as sp
    d = copy.deepcopy(get_desc("ur10"))
    d.base_t = np.array([0.1, 0.0, 0.0], np.float32)
    with pytest.raises(ValueError, match="base"):
        sp.specialized_header(d)
    d = get_desc("ur10")
    with pytest.raises(ValueError, match="<= 16 points"):
        sp.specialized_header(d, dict(pt_link=[1] * 17, pairs=[], ee_body=7))


def test_private_library_builds_loads_and_exports_the_abi():
    """nvcc cross-compiles for sm_100a without a GPU; the result must export every entry point of include/trk_b200.h."""
    from torch_robotics_b200 import _lib
    for path in sc.build_all():
        assert os.path.exists(path)
        L = _lib.load_library(path)                     # raises when an export of _lib.EXPORTS is missing
        assert L.trk_abi_version() == 1
        assert L is _lib.load_library(path)             # one handle per path


def test_tree_specialize_bookkeeping_without_a_gpu():
    """DifferentiableTree.specialize builds (here: finds) the library, bumps the model generation so that dependants rebuild
    their device costs, and is dropped again when the tree changes; a collision model needs the EE link."""
    from torch_robotics_b200 import _lib
    from torch_robotics_b200.torch_kinematics_tree.models.robot_tree import DifferentiableTree
    tree = DifferentiableTree(os.path.join(sc.DATA, "kuka_iiwa", "urdf", "iiwa7_allegro.urdf"), "iiwa7_allegro", device="cuda:0")
    assert tree._library_for_current_tree() is None
    gen = tree._model_generation
    path = tree.specialize()
    assert os.path.exists(path) and tree._library_for_current_tree() == path and tree._model_generation == gen + 1
    assert path == sc.build_all()[0]                                   # the content-addressed cache: the same robot, the same file
    tree.attach_points([("palm_link", (0.0, 0.0, 0.05))])              # another tree: the default library serves it again
    assert tree._library_for_current_tree() is None
    with pytest.raises(ValueError, match="ee_link"):
        tree.specialize(collision_points=["palm_link"])
    with pytest.raises(ImportError, match="not found"):
        _lib.load_library(os.path.join(os.path.dirname(path), "libtrk_jit_does_not_exist.so"))


# --------------------------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")
DEV = "cuda:0"


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


@pytest.mark.gpu
def test_gpu_specialised_fk_matches_oracle_and_runtime_tree_kernels():
    from oracle import restatement as orc
    from torch_robotics_b200 import specialize as sp
    from torch_robotics_b200._lib import DeviceModel
    d = sc.iiwa7_allegro_desc()
    generic = DeviceModel(d, DEV)
    special = DeviceModel(d, DEV, library=sp.build_specialized_library(d))
    assert generic.kernel_family == "generic" and special.kernel_family == "iiwa7_allegro"
    rng = np.random.default_rng(5)
    q = rng.uniform(-3.0, 3.0, (3001, d.n_dofs)).astype(np.float32)            # beyond several joint limits: clamped joints
    Hs, Hg = special.fk_forward(t(q)), generic.fk_forward(t(q))
    Rw, pw = orc.fk_forward(d, q, np.float32)
    assert rel_err(Hs.cpu().numpy(), orc.homogeneous(Rw, pw)) < POSE_TOL
    assert rel_err(Hs.cpu().numpy(), Hg.cpu().numpy()) < POSE_TOL
    G = rng.standard_normal(Hs.shape).astype(np.float32)
    gs, gg = special.fk_backward(t(q), t(G)), generic.fk_backward(t(q), t(G))
    go = orc.fk_backward(d, q, G, dtype=np.float64)
    assert rel_err(gs.cpu().numpy(), go) < GRAD_TOL
    assert rel_err(gs.cpu().numpy(), gg.cpu().numpy()) < GRAD_TOL
    # sub-selection of links (the staged store path) and the link Jacobians
    sel = [28, 3, 11]
    assert torch.equal(special.fk_forward(t(q), sel), Hs[:, sel])
    Js, Jg = special.link_jacobian(t(q[:257])), generic.link_jacobian(t(q[:257]))
    assert rel_err(Js.cpu().numpy(), Jg.cpu().numpy()) < GRAD_TOL


@pytest.mark.gpu
def test_gpu_specialised_collision_model_runs_the_headline_kernel_and_matches_the_oracle():
    """UR10 with an ad-hoc collision model: the default library serves it with the table-driven kernel, the specialised one
    with k_objective_hot<Jit, JitCollision> (the fused block sums only exist there).  Both against the C oracle."""
    from oracle import c_oracle as co
    from torch_robotics_b200 import specialize as sp
    from torch_robotics_b200._lib import DeviceCost, DeviceModel
    from torch_robotics_b200.descriptor import build_cost_descriptor
    d, cm = sc.ur10_collision()
    special = DeviceModel(d, DEV, library=sp.build_specialized_library(d, cm))
    default = DeviceModel(d, DEV)
    assert special.kernel_family == "ur10"
    rng = np.random.default_rng(11)
    centers = rng.uniform(-0.7, 0.7, (10, 3)).astype(np.float32)
    target = np.eye(4, dtype=np.float32)
    target[:3, 3] = [0.4, 0.2, 0.5]
    cost = build_cost_descriptor(pt_link=cm["pt_link"], pt_radius=rng.uniform(0.04, 0.1, len(cm["pt_link"])).astype(np.float32), cutoff_obj=0.02,
                                 objects=[(centers, np.full(10, 0.12, np.float32), None, None)], ws_min=[-1.2, -1.2, -0.4], ws_max=[1.2, 1.2, 1.4],
                                 self_pairs=cm["pairs"], self_margin=[0.05], ee_link=cm["ee_body"], ee_target=target)
    q = rng.uniform(-np.pi, np.pi, (20_003, d.n_dofs)).astype(np.float32)
    dcs, dcd = DeviceCost(special, cost), DeviceCost(default, cost)
    r = dcs.objective(t(q), want_terms=True, want_occupancy=True, want_argmin=True)
    r0 = dcd.objective(t(q), want_terms=True, want_occupancy=True, want_argmin=True)
    o = co.objective(co.Tree(d), co.Cost(cost), q)
    assert rel_err(r["terms"].cpu().numpy(), o["terms"]) < POSE_TOL
    assert rel_err(r["total"].cpu().numpy(), r0["total"].cpu().numpy()) < POSE_TOL
    gd = np.abs(r["grad_q"].cpu().numpy() - o["grad_q"]).max(axis=1) / np.abs(o["grad_q"]).max()
    assert np.mean(gd < GRAD_TOL) > 0.999 and gd[gd < GRAD_TOL].max() < GRAD_TOL        # the rest: arg-min near ties (checked below)
    same_argmin = np.all((r["argmin"].cpu().numpy() == o["argmin"]).reshape(len(q), -1), axis=1)
    assert np.all(gd[same_argmin] < GRAD_TOL)
    assert np.mean(same_argmin) > 0.999 and np.mean(r["occupancy"].cpu().numpy() == o["occ"]) > 0.999
    # only the headline kernel has the fused first stage of the cost reduction
    total, grad = torch.empty(len(q), device=DEV), torch.empty(len(q), d.n_dofs, device=DEV)
    parts = dcs.objective_sum(t(q), out_grad=grad, out_total=total)
    assert abs(float(parts.sum(dtype=torch.float64)) - float(total.sum(dtype=torch.float64))) < 1e-3 * abs(float(total.sum(dtype=torch.float64)))
    with pytest.raises(RuntimeError, match="fused cost reduction"):
        dcd.objective_sum(t(q), out_grad=grad, out_total=total)


@pytest.mark.gpu
def test_gpu_tree_specialize_switches_the_device_model():
    from torch_robotics_b200.torch_kinematics_tree.models.robot_tree import DifferentiableTree
    tree = DifferentiableTree(os.path.join(sc.DATA, "kuka_iiwa", "urdf", "iiwa7_allegro.urdf"), "iiwa7_allegro", device=DEV)
    q = torch.rand(64, tree._n_dofs, device=DEV) - 0.5
    H0 = tree.compute_forward_kinematics_all_links(q)
    assert tree.device_model(DEV).kernel_family == "generic"
    tree.specialize()
    assert tree.device_model(DEV).kernel_family == "iiwa7_allegro"
    H1 = tree.compute_forward_kinematics_all_links(q)
    assert rel_err(H1.cpu().numpy(), H0.cpu().numpy()) < POSE_TOL
    tree.attach_points([("palm_link", (0.0, 0.0, 0.05))])        # another tree: back to the default library
    assert tree.device_model(DEV).kernel_family == "generic"
