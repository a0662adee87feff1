"""Named robot models bound to their URDF (reference: models/robots.py:16-133)."""
from typing import Optional

from ..utils.files import get_robot_path
from .robot_tree import DifferentiableTree


def _make(cls_name, rel_path, model_name, doc):
    def __init__(self, link_list: Optional[str] = None, device="cpu"):
        self.model_path = (get_robot_path().joinpath(*rel_path)).as_posix()
        DifferentiableTree.__init__(self, self.model_path, model_name, link_list=link_list, device=device)
    return type(cls_name, (DifferentiableTree,), {"__init__": __init__, "__doc__": doc})


DifferentiableKUKAiiwa = _make("DifferentiableKUKAiiwa", ("kuka_iiwa", "urdf", "iiwa7.urdf"),
                               "differentiable_kuka_iiwa", "KUKA iiwa7 (robots.py:16)")
DifferentiableHabitatStretch = _make("DifferentiableHabitatStretch", ("habitat_stretch", "urdf", "hab_stretch.urdf"),
                                     "differentiable_stretch", "Habitat Stretch, 44 links / 14 DoF (robots.py:84)")
DifferentiableTiagoDualHolo = _make("DifferentiableTiagoDualHolo",
                                    ("tiago_dual_description", "tiago_dual_holobase_minimal.urdf"),
                                    "differentiable_tiago_dual_holo", "Tiago dual-arm (robots.py:93)")
DifferentiableShadowHand = _make("DifferentiableShadowHand", ("shadow_hand", "shadow_hand.urdf"),
                                 "differentiable_shadow_hand", "Shadow hand (robots.py:112)")
DifferentiableAllegroHand = _make("DifferentiableAllegroHand", ("allegro_hand", "allegro_hand.urdf"),
                                  "differentiable_allegro_hand", "Allegro hand (robots.py:120)")
Differentiable2LinkPlanar = _make("Differentiable2LinkPlanar", ("planar_manipulators", "urdf", "2_link_planar.urdf"),
                                  "differentiable_2_link_planar", "2-link planar arm (robots.py:128)")


class DifferentiableFrankaPanda(DifferentiableTree):
    """Franka Panda with or without gripper (robots.py:56-69)."""

    def __init__(self, link_list: Optional[str] = None, gripper=False, device="cpu", grasped_object=None):
        fname = "panda_arm_hand.urdf" if gripper else "panda_arm_no_gripper.urdf"
        self.model_path = (get_robot_path() / "franka_description" / "robots" / fname).as_posix()
        super().__init__(self.model_path, "differentiable_franka_panda", link_list=link_list, device=device)
        if grasped_object is not None:
            # The reference writes a second URDF with a fixed joint panda_hand -> grasped_object whose origin is
            # (pos, q_to_euler(ori)) printed as fp32 (robots.py:24-53); the same link is appended in memory here.
            from ..geometrics.quaternion import q_to_euler
            xyz = grasped_object.pos.detach().float().cpu().reshape(-1).tolist()
            rpy = q_to_euler(grasped_object.ori.detach().float().cpu()).reshape(-1).tolist()
            self.attach_link("grasped_object", "panda_hand", xyz, rpy)


class DifferentiableUR10(DifferentiableTree):
    def __init__(self, link_list: Optional[str] = None, attach_gripper=False, device="cpu"):
        fname = "ur10_suction.urdf" if attach_gripper else "ur10.urdf"
        self.model_path = (get_robot_path() / "ur10" / "urdf" / fname).as_posix()
        super().__init__(self.model_path, "differentiable_ur10", link_list=link_list, device=device)


class DifferentiableTiagoDualHoloMove(DifferentiableTree):
    def __init__(self, link_list: Optional[str] = None, device="cpu"):
        self.model_path = (get_robot_path() / "tiago_dual_description" /
                           "tiago_dual_holobase_minimal_holonomic.urdf").as_posix()
        super().__init__(self.model_path, "differentiable_tiago_dual_holo_move", link_list=link_list, device=device)

    def get_link_names(self):
        # the three virtual base links are hidden from callers (robots.py:108)
        return super().get_link_names()[3:]
