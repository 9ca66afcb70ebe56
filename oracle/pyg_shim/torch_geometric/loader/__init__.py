"""DataLoader(dataset, batch_size, shuffle=False): mini-batches of HeteroData via Batch.from_data_list
(test_amazon_filterd.py:488,547; fine_tune_ours.py:790,814)."""
import torch

from ..data import Batch


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)
