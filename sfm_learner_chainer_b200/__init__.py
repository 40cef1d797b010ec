"""B200-native view-synthesis loss path of SfM-Learner (drop-in for pfnet/sfm-learner-chainer).

Host-side mirror of the reference interface for this path:
  models/transform.py:156            projective_inverse_warp          -> functions.projective_inverse_warp
  models/spational_transformer_sampler_interp.py:152                  -> functions.spatial_transformer_sampler_interp
  models/base_model.py:48-124        SFMLearner.__call__ loss loop    -> functions.ViewSynthesisLoss / base_model.SFMLearner
All arithmetic runs in hand-written sm_100a kernels behind the C ABI of include/sfmloss.h.
"""
from . import lib
from .functions import (ViewSynthesisLoss, projective_inverse_warp, projective_inverse_warp_backward,
                        spatial_transformer_sampler_interp, SpatialTransformerSamplerInterp,
                        disp_activation, pose_reduce, ingest_u8, draw_augmentation, evaluate_depth_batch)
from .base_model import SFMLearner

__all__ = ['lib', 'ViewSynthesisLoss', 'projective_inverse_warp', 'projective_inverse_warp_backward',
           'spatial_transformer_sampler_interp', 'SpatialTransformerSamplerInterp', 'SFMLearner',
           'disp_activation', 'pose_reduce', 'ingest_u8', 'draw_augmentation', 'evaluate_depth_batch']
