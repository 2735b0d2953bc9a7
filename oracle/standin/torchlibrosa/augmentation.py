"""ORACLE / TEST INFRASTRUCTURE ONLY -- ``SpecAugmentation`` is only called when
``self.training`` (reference ``mellow/model/htsat.py:871-872``), so an identity
module with the same constructor signature is sufficient for the inference path."""
import torch.nn as nn


class SpecAugmentation(nn.Module):
    def __init__(self, time_drop_width, time_stripes_num, freq_drop_width, freq_stripes_num):
        super().__init__()

    def forward(self, x):
        return x
