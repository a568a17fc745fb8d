"""The two helpers of the reference's utils/utils.py that the head path needs."""
import random

import numpy as np
import torch


class Struct:
    """utils/utils.py:246-248 -- attribute bag built from the YAML/argparse dict."""

    def __init__(self, **entries):
        self.__dict__.update(entries)


def set_seed(seed: int) -> None:
    """utils/utils.py:226-244."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    if torch.backends.cudnn.enabled:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
