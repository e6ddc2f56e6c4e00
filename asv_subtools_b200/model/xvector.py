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

from asv_subtools_b200 import ops  # noqa: E402
from asv_subtools_b200.nnet import (ReluBatchNormTdnnLayer, StatisticsPooling,  # noqa: E402
                                    TopVirtualNnet)


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
        if self.extracted_embedding not in ("far", "near"):
            raise TypeError("Expected far or near position, but got {}".format(self.extracted_embedding))
        self.device_for_extraction()
        ex = ops.Extractor(self.inputs_dim)

        def arrays(layer):
            w = layer.affine.weight.detach().float().cpu().numpy()
            b = layer.affine.bias.detach().float().cpu().numpy() if layer.affine.bias is not None else None
            scale, shift = layer.folded_bn()
            return w, b, scale, shift

        for layer in (self.tdnn1, self.tdnn2, self.tdnn3, self.tdnn4, self.tdnn5):
            w, b, scale, shift = arrays(layer)
            ex.add_frame_layer(w, b, layer.affine.context, scale, shift, relu=layer.relu)
        w, b, scale, shift = arrays(self.tdnn6)
        if self.extracted_embedding == "far":       # xvector.py:92-93: tdnn6.affine only
            ex.add_segment_layer(w, b)
        else:                                       # xvector.py:94-96: tdnn6 (full) -> tdnn7.affine
            ex.add_segment_layer(w, b, scale, shift, relu=self.tdnn6.relu)
            w7, b7, _, _ = arrays(self.tdnn7)
            ex.add_segment_layer(w7, b7)
        ex.finalize(pooling_eps=self.stats.eps)
        return ex


if __name__ == "__main__":
    print(Xvector(23, 1211))
