"""Drop-in ``rnampnn.utils.seed.seeding`` (reference: rnampnn/utils/seed.py:9-19): seeds every RNG the package
touches and asks for 'medium' fp32 matmul precision like the reference does."""
import os
import random

import numpy as np
import torch

from ..config.glob import DEFAULT_SEED


def seeding(seed=DEFAULT_SEED):
    os.environ['PYTHONHASHSEED'] = str(seed)
    for seeder in (random.seed, np.random.seed, torch.manual_seed, torch.cuda.manual_seed, torch.cuda.manual_seed_all):
        seeder(seed)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.set_float32_matmul_precision('medium')
    print('seeding done!!!')
