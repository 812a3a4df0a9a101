"""dmvae_b200: B200-native (sm_100a) implementation of the DMVAE data-parallel training hot path.

Public surface mirrors the reference's modules for this path:
    dmvae_b200.autoencoder  <->  models/flux_ae.py   (Encoder, Decoder, ResnetBlock, AttnBlock, Upsample, Downsample)
    dmvae_b200.vae          <->  models/vae.py       (VAE, DINOEncoder, MLP)
    dmvae_b200.lpips        <->  utils/lpips.py      (LPIPS)
    dmvae_b200.train        <->  VAELossFunction / step logic of train_dmd.py, train_tokenizer.py
    dmvae_b200.losses       fused DMD / L1+L2 / LPIPS-distance / reparam+KL operators
All arithmetic runs in libdmvae_b200.so (include/dmvae_b200.h); there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401
from ._lib import DmvaeError, LIB_PATH  # noqa: F401

__all__ = ["DmvaeError", "LIB_PATH"]
