"""StatisticsPooling mirror (pytorch/libs/nnet/pooling.py:15-76).  No parameters; arithmetic in
csrc/pooling.cu."""
import torch

from .. import ops


class StatisticsPooling(torch.nn.Module):
    def __init__(self, input_dim, stddev=True, unbiased=False, eps=1.0e-10):
        super().__init__()
        if not stddev or unbiased:
            raise NotImplementedError("B200 StatisticsPooling implements the mean+std, biased-variance case")
        self.input_dim, self.stddev, self.unbiased, self.eps = input_dim, stddev, unbiased, eps
        self.output_dim = 2 * input_dim

    def get_output_dim(self):
        return self.output_dim

    def forward(self, inputs):
        """inputs: (B, C, T) like the reference -> (B, 2C, 1)."""
        x = inputs.transpose(1, 2).contiguous().float()
        return ops.stats_pool(x, eps=self.eps).unsqueeze(2)
