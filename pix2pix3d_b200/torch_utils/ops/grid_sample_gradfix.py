"""`grid_sample` entry point kept for surface compatibility (reference torch_utils/ops/grid_sample_gradfix.py).

The switch is off everywhere on the hot path (training_loop.py:282) and the renderer never calls it, so
this simply forwards to the 2-D bilinear/zeros/align_corners=False configuration the reference pins."""
import torch

enabled = False


def grid_sample(input, grid):
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros',
                                           align_corners=False)
