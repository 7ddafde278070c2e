from typing import Union

import numpy as np
import torch

ArrayLike = Union[torch.Tensor, np.ndarray, np.generic, bool, int, float, complex]
DTypeLike = object
