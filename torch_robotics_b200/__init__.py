"""B200-native drop-in for the FK + planning-objective hot path of anindex/torch_robotics.

Importing the package does not need a GPU; every compute entry point does, and raises otherwise.
"""
from .descriptor import CostDescriptor, TreeDescriptor, build_cost_descriptor, compile_urdf  # noqa: F401

__all__ = ["CostDescriptor", "TreeDescriptor", "build_cost_descriptor", "compile_urdf"]
