"""Model-level boundary: the reference's class names (mmdet3d/models/ReIDNet.py:40-96 `module_obj`)."""
from .attention import corss_attention, cross_lin_attn, local_self_attention
from .backbone_net import Pointnet_Backbone
from .builder import FUSIONMODELS, build_fusion_model, build_model
from .dgcnn_orig import DGCNN
from .lanegcn_nets import LinearRes
from .pointnet import PointNet, PointNetEncoder, STN3d, STNkd
from .pointnet2_utils import (FP_SA, LinearAttention, PointNetFeaturePropagationSA, PointNetSetAbstractionEdgeSA,
                              Self_Attention)
from .ReIDNet import ReIDNet, build_module, build_sequential, module_obj
from .image_reid import ImageReIDNet

__all__ = ["ReIDNet", "ImageReIDNet", "cross_lin_attn", "local_self_attention", "Pointnet_Backbone", "DGCNN", "PointNet", "corss_attention", "LinearRes", "FUSIONMODELS",
           "build_model", "build_fusion_model", "build_module", "build_sequential", "module_obj"]
