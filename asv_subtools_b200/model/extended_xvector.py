# -*- coding:utf-8 -*-
"""Extended x-vector blueprint for the B200 path -- drop-in for pytorch/model/extended_xvector.py
(ExtendedXvector.init :15-49, extract_embedding :93-116): the standard TDNN stack with 1x1 layers
interleaved and an extra [-4,0,4] layer.  Same constructor signature and state_dict keys; it reuses the
native frame-layer / fused-pooling / segment-layer extractor unchanged."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from asv_subtools_b200.nnet import (ReluBatchNormTdnnLayer, StatisticsPooling,  # noqa: E402
                                    TopVirtualNnet, build_tdnn_extractor)


class ExtendedXvector(TopVirtualNnet):
    """An extended x-vector framework."""

    def init(self, inputs_dim, num_targets, extend=True, nonlinearity="relu", aug_dropout=0.2, training=True,
             extracted_embedding="far"):
        if nonlinearity != "relu":
            raise NotImplementedError("B200 path implements the reference default nonlinearity='relu'")
        self.inputs_dim = inputs_dim
        self.extracted_embedding = extracted_embedding
        L = ReluBatchNormTdnnLayer
        self.tdnn1 = L(inputs_dim, 512, [-2, -1, 0, 1, 2], nonlinearity=nonlinearity)
        self.ex_tdnn1 = L(512, 512, nonlinearity=nonlinearity) if extend else None
        self.tdnn2 = L(512, 512, [-2, 0, 2], nonlinearity=nonlinearity)
        self.ex_tdnn2 = L(512, 512, nonlinearity=nonlinearity) if extend else None
        self.tdnn3 = L(512, 512, [-3, 0, 3], nonlinearity=nonlinearity)
        self.ex_tdnn3 = L(512, 512, nonlinearity=nonlinearity) if extend else None
        self.ex_tdnn4 = L(512, 512, [-4, 0, 4], nonlinearity=nonlinearity) if extend else None
        self.ex_tdnn5 = L(512, 512, nonlinearity=nonlinearity) if extend else None
        self.tdnn4 = L(512, 512, nonlinearity=nonlinearity)
        self.tdnn5 = L(512, 1500, nonlinearity=nonlinearity)
        self.stats = StatisticsPooling(1500, stddev=True)
        self.tdnn6 = L(self.stats.get_output_dim(), 512, nonlinearity=nonlinearity)
        self.tdnn7 = L(512, 512, nonlinearity=nonlinearity)
        self.transform_keys = ["tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5", "stats", "tdnn6", "tdnn7",
                               "ex_tdnn1", "ex_tdnn2", "ex_tdnn3", "ex_tdnn4", "ex_tdnn5"]

    def build_extractor(self):
        order = (self.tdnn1, self.ex_tdnn1, self.tdnn2, self.ex_tdnn2, self.tdnn3, self.ex_tdnn3, self.ex_tdnn4,
                 self.ex_tdnn5, self.tdnn4, self.tdnn5)
        return build_tdnn_extractor(self, self.inputs_dim, [l for l in order if l is not None], self.stats, self.tdnn6,
                                    self.tdnn7, self.extracted_embedding)


if __name__ == "__main__":
    print(ExtendedXvector(23, 1211))
