"""placeholder: the reference scripts import torch_geometric.transforms as T but the encoder path never
calls it."""
