"""FK of all links for the shipped robots - the B200 counterpart of the reference's
examples/forward_kinematics.py (same calls; device is CUDA because this implementation has no CPU path)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_robotics_b200.torch_kinematics_tree.models.robots import (  # noqa: E402
    DifferentiableAllegroHand, DifferentiableFrankaPanda, DifferentiableHabitatStretch, DifferentiableShadowHand,
    DifferentiableTiagoDualHoloMove, DifferentiableUR10)
from torch_robotics_b200.torch_utils.seed import fix_random_seed  # noqa: E402
from torch_robotics_b200.torch_utils.torch_timer import TimerCUDA  # noqa: E402

if __name__ == "__main__":
    fix_random_seed(1)
    batch_size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    device = "cuda:0"
    for title, cls in [("Panda", DifferentiableFrankaPanda), ("UR10", DifferentiableUR10),
                       ("Habitat Stretch", DifferentiableHabitatStretch), ("Tiago", DifferentiableTiagoDualHoloMove),
                       ("Shadow Hand", DifferentiableShadowHand), ("Allegro Hand", DifferentiableAllegroHand)]:
        print(f"\n=========================== {title} Model ===============================")
        model = cls(device=device)
        print(model._n_dofs, "dofs;", len(model.get_link_names()), "links; kernel family:", model.device_model(device).kernel_family)
        q = torch.rand(batch_size, model._n_dofs).to(device).requires_grad_(True)
        model.compute_forward_kinematics_all_links(q)          # warm-up (table upload)
        with TimerCUDA() as t:
            data = model.compute_forward_kinematics_all_links(q)
        print(tuple(data.shape), f"Computational Time {t.elapsed:.6f}")
