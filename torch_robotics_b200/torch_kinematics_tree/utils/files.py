"""Data-path helpers (reference: torch_kinematics_tree/utils/files.py:7-65).  The package ships
kinematics-only URDF skeletons under torch_robotics_b200/data (see tools/strip_urdf.py); set
TORCH_ROBOTICS_DATA to point at a full torch_robotics data directory instead."""
import os
from pathlib import Path

import yaml


def get_root_path():
    return Path(__file__).resolve().parents[2]


def get_data_path():
    env = os.environ.get("TORCH_ROBOTICS_DATA")
    return Path(env) if env else get_root_path() / "data"


def get_urdf_path():
    return get_data_path() / "urdf"


def get_robot_path():
    return get_urdf_path() / "robots"


def get_objects_path():
    return get_urdf_path() / "objects"


def get_configs_path():
    return get_data_path() / "configs"


def load_yaml(filename):
    with open(filename, "r") as stream:
        return yaml.safe_load(stream)
