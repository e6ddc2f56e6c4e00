# -*- coding:utf-8 -*-
"""Standard x-vector blueprint for the B200 path -- drop-in for pytorch/model/xvector.py.

Point `model_blueprint` in `<model_dir>/config/nnet.config` at this file and keep the creation
string (`Xvector(23,1211,training=False,extracted_embedding="far",...)`): the constructor
signature, the state_dict keys (`tdnnK.affine.weight` (Cout,Cin,tot_context) with masked taps,
`.affine.bias`, `.batchnorm.{weight,bias,running_mean,running_var,num_batches_tracked}`) and
`extract_embedding()` are the reference's (xvector.py:18-40, :77-98); the arithmetic is CUDA.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from asv_subtools_b200.nnet import (ReluBatchNormTdnnLayer, StatisticsPooling,  # noqa: E402
                                    TopVirtualNnet, build_tdnn_extractor)


class Xvector(TopVirtualNnet):
    """A standard x-vector framework (tdnn1-5 -> stats -> tdnn6 [-> tdnn7])."""

    def init(self, inputs_dim, num_targets, nonlinearity="relu", aug_dropout=0.2, training=True,
             extracted_embedding="far"):
        if nonlinearity != "relu":
            raise NotImplementedError("B200 x-vector path implements the reference default nonlinearity='relu'")
        self.inputs_dim = inputs_dim
        self.extracted_embedding = extracted_embedding
        self.tdnn1 = ReluBatchNormTdnnLayer(inputs_dim, 512, [-2, -1, 0, 1, 2], nonlinearity=nonlinearity)
        self.tdnn2 = ReluBatchNormTdnnLayer(512, 512, [-2, 0, 2], nonlinearity=nonlinearity)
        self.tdnn3 = ReluBatchNormTdnnLayer(512, 512, [-3, 0, 3], nonlinearity=nonlinearity)
        self.tdnn4 = ReluBatchNormTdnnLayer(512, 512, nonlinearity=nonlinearity)
        self.tdnn5 = ReluBatchNormTdnnLayer(512, 1500, nonlinearity=nonlinearity)
        self.stats = StatisticsPooling(1500, stddev=True)
        self.tdnn6 = ReluBatchNormTdnnLayer(self.stats.get_output_dim(), 512, nonlinearity=nonlinearity)
        self.tdnn7 = ReluBatchNormTdnnLayer(512, 512, nonlinearity=nonlinearity)
        # Training-only pieces (aug_dropout, SoftmaxLoss) are outside the extraction path; loss.*
        # keys of a training checkpoint are ignored by load_state_dict(strict=False) as in the
        # reference's extract_embeddings.py:63.
        self.transform_keys = ["tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5", "stats", "tdnn6", "tdnn7"]

    def build_extractor(self):
        return build_tdnn_extractor(self, self.inputs_dim, (self.tdnn1, self.tdnn2, self.tdnn3, self.tdnn4, self.tdnn5),
                                    self.stats, self.tdnn6, self.tdnn7, self.extracted_embedding)


if __name__ == "__main__":
    print(Xvector(23, 1211))
