"""spatial_vae for NVIDIA B200: drop-in for the reference package's training-step hot path.

`spatial_vae.models` keeps the reference module API; the arithmetic runs in libsvae_b200.so
(hand-written sm_100a CUDA behind the C ABI of include/svae_b200.h).
"""
