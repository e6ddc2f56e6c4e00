# -*- coding:utf-8 -*-
"""Snowdar x-vector blueprint for the B200 path -- drop-in for pytorch/model/snowdar_xvector.py (Xvector.init
:15-152, extract_embedding :262-294) in its TDNN configurations: standard or `extend=True` stack,
`tdnn_layer_params` (default BatchNorm affine=False, momentum 0.5), pooling = statistics | attentive | multi-head |
multi-resolution | lde | xi-postmean-softplus2 | xi-postdist-softplus2 (pooling.py:15-76, :130-212, :322-440, :518-587; the
blueprint's switch :119-136 in full), positions
far / near_affine / near.  Same constructor keywords and state_dict keys.  The options that add other
operators (SE blocks, skip connection) raise NotImplementedError;
training-only keywords (mixup, specaugment, dropouts, margin loss, step params) are accepted and ignored,
as the launchers rewrite the creation string with training=False for extraction."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from asv_subtools_b200.nnet import (AttentionPoolingExtractor, AttentiveStatisticsPooling, LDEPooling,  # noqa: E402
                                    MultiHeadAttentionPooling, MultiResolutionMultiHeadAttentionPooling,
                                    xivec_stdinit_softplus2_prec_pooling,
                                    ReluBatchNormTdnnLayer, StatisticsPooling, TopVirtualNnet, build_tdnn_extractor)


class Xvector(TopVirtualNnet):
    """A composite x-vector framework."""

    def init(self, inputs_dim, num_targets, extend=False, skip_connection=False, mixup=False, mixup_alpha=1.0,
             specaugment=False, specaugment_params={}, aug_dropout=0., context_dropout=0., hidden_dropout=0.,
             dropout_params={}, SE=False, se_ratio=4, tdnn_layer_params={}, tdnn6=True, tdnn7_params={},
             pooling="statistics", pooling_params={}, margin_loss=False, margin_loss_params={}, use_step=False,
             step_params={}, transfer_from="softmax_loss", training=True, extracted_embedding="far"):
        if SE or skip_connection:
            raise NotImplementedError("SE blocks / skip connection are not on the B200 path")
        if pooling not in ("statistics", "lde", "attentive", "multi-head", "multi-resolution", "xi-postmean-softplus2",
                           "xi-postdist-softplus2"):
            raise NotImplementedError("pooling={!r} is not one of the reference's options (snowdar_xvector.py:119-136)".format(pooling))
        if not tdnn6:
            raise NotImplementedError("tdnn6=False is not on the B200 path")
        layer = {"nonlinearity": "relu", "nonlinearity_params": {"inplace": True}, "bn-relu": False, "bn": True,
                 "bn_params": {"momentum": 0.5, "affine": False, "track_running_stats": True}}     # :45-48
        layer.update(tdnn_layer_params)
        last = dict(layer)
        last.update(tdnn7_params)
        if last.get("nonlinearity") == "default":
            last["nonlinearity"] = layer["nonlinearity"]
        pool = {"num_nodes": 1500, "num_head": 1, "share": True, "affine_layers": 1, "hidden_size": 64, "context": [0],
                "stddev": True, "temperature": False, "fixed": True}                                      # :44-55
        pool.update(pooling_params)
        num_nodes = pool.pop("num_nodes")
        if not pool.pop("stddev"):
            raise NotImplementedError("stddev=False is not on the B200 path")
        self.inputs_dim = inputs_dim
        self.extracted_embedding = extracted_embedding
        L = ReluBatchNormTdnnLayer
        self.tdnn1 = L(inputs_dim, 512, [-2, -1, 0, 1, 2], **layer)
        self.ex_tdnn1 = L(512, 512, **layer) if extend else None
        self.tdnn2 = L(512, 512, [-2, 0, 2], **layer)
        self.ex_tdnn2 = L(512, 512, **layer) if extend else None
        self.tdnn3 = L(512, 512, [-3, 0, 3], **layer)
        self.ex_tdnn3 = L(512, 512, **layer) if extend else None
        self.ex_tdnn4 = L(512, 512, [-4, 0, 4], **layer) if extend else None
        self.ex_tdnn5 = L(512, 512, **layer) if extend else None
        self.tdnn4 = L(512, 512, **layer)
        self.tdnn5 = L(512, num_nodes, **layer)
        if pooling == "lde":                                                                                # :121-122
            self.stats = LDEPooling(num_nodes, c_num=pool["num_head"])
        elif pooling == "attentive":                                                                        # :123-126
            self.stats = AttentiveStatisticsPooling(num_nodes, affine_layers=pool["affine_layers"],
                                                    hidden_size=pool["hidden_size"], context=pool["context"], stddev=True)
        elif pooling == "multi-head":                                                                       # :127-128
            self.stats = MultiHeadAttentionPooling(num_nodes, stddev=True, **pool)
        elif pooling == "multi-resolution":                                                                 # :129-130
            self.stats = MultiResolutionMultiHeadAttentionPooling(num_nodes, **pool)
        elif pooling in ("xi-postmean-softplus2", "xi-postdist-softplus2"):                                 # :131-134
            self.stats = xivec_stdinit_softplus2_prec_pooling(num_nodes, hidden_size=pool["hidden_size"],
                                                              stddev=pooling == "xi-postdist-softplus2")
        else:
            self.stats = StatisticsPooling(num_nodes, stddev=True)
        self.tdnn6 = L(self.stats.get_output_dim(), 512, **layer)
        self.tdnn7 = L(512, 512, **last)
        self.transform_keys = ["tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5", "stats", "tdnn6", "tdnn7", "ex_tdnn1",
                               "ex_tdnn2", "ex_tdnn3", "ex_tdnn4", "ex_tdnn5", "se1", "se2", "se3", "se4", "loss"]

    def build_extractor(self):
        order = (self.tdnn1, self.ex_tdnn1, self.tdnn2, self.ex_tdnn2, self.tdnn3, self.ex_tdnn3, self.ex_tdnn4,
                 self.ex_tdnn5, self.tdnn4, self.tdnn5)
        pos = {"far": "far", "near_affine": "near_affine", "near": "near_full"}.get(self.extracted_embedding)
        if pos is None:
            raise TypeError("Expected far or near position, but got {}".format(self.extracted_embedding))
        layers = [l for l in order if l is not None]
        if not isinstance(self.stats, StatisticsPooling):
            return AttentionPoolingExtractor(self, self.inputs_dim, layers, self.stats, self.tdnn6, self.tdnn7, pos)
        return build_tdnn_extractor(self, self.inputs_dim, layers, self.stats, self.tdnn6, self.tdnn7, pos)


if __name__ == "__main__":
    print(Xvector(23, 1211))
