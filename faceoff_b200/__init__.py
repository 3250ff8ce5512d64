"""faceoff_b200 -- B200-native (sm_100a) implementation of FaceOff's VQVAE-conv3d (+LPIPS) training-step hot path.

Drop-in replacements for the reference's nn.Modules (models/vqvae_conv3d_latent.py, models/lpips.py,
loss.py:VQLPIPS) and its ``distributed`` package, over a C-ABI CUDA library (include/faceoff_b200.h).
"""
__version__ = "0.1.0"
