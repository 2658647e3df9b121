"""``set_seed`` and ``acc`` (gnntf/measures.py:7-14)."""
import random

import numpy as np
import torch


def set_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def acc(predictions, labels):
    predictions, labels = _np(predictions), _np(labels)
    return 1 - np.count_nonzero(predictions - labels) / predictions.shape[0]
