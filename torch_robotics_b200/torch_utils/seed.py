"""fix_random_seed (reference: torch_utils/seed.py:7)."""
import random

import numpy as np
import torch


def fix_random_seed(seed):
    """Seed python, numpy and torch (CPU and CUDA generators); cuDNN autotuning off so that runs repeat."""
    seed = int(seed)
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
