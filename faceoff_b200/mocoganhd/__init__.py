"""MoCoGAN-HD discriminators of the FaceOff discriminator trainer (BASELINE configs[4]):
drop-in replacements for TemporalAlignment/models/mocoganhd_content_disc.py, mocoganhd_video_disc.py and the GAN losses of
mocoganhd_losses.py that disc_trainers/train_vqvae_perceptual_mocoganhd_disc.py:160-333 uses.  Same class names, constructor
arguments, forward results (lists of per-scale feature lists) and state-dict keys; every tensor operation runs in the CUDA
kernels of csrc/disc.cu behind the C ABI (no PyTorch / CPU fallback).

    from faceoff_b200.mocoganhd import content_disc as mocoganhd_content_disc
    from faceoff_b200.mocoganhd import video_disc as mocoganhd_video_disc
    from faceoff_b200.mocoganhd import losses as mocoganhd_losses
"""
from . import content_disc, losses, video_disc  # noqa: F401
