"""hunyuanvideo_efficiency_b200 — B200-native (sm_100a) HunyuanVideo 3D causal VAE encode/decode path.

Drop-in for `hyvideo.vae` of c976237222/HunyuanVideo_efficiency:

    from hunyuanvideo_efficiency_b200.vae import load_vae, AutoencoderKLCausal3D

All compute runs in hand-written CUDA kernels behind a C ABI (include/hyvae.h, libhyvae.so).
"""
__version__ = "0.1.0"
