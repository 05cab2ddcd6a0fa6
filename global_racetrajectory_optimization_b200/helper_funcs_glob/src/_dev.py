"""numpy <-> device helpers shared by the single-track mirrors (batch of one through the C-ABI)."""
import numpy as np
import torch

from ... import batch as _b


def device():
    _b._require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def up(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device()).unsqueeze(0)
