import torch

_COLORS = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0), "red": (1.0, 0.0, 0.0), "green": (0.0, 1.0, 0.0),
           "blue": (0.0, 0.0, 1.0)}


def get_color(color):
    if isinstance(color, str):
        return torch.tensor(_COLORS[color.lower()])
    return torch.tensor(color)
