"""Architecture definitions (plain PyTorch, random init) of the model families the
engine knows: FLUX.1 MM-DiT, SD1.5/SDXL UNet, WAN2.x video DiT, Z-Image (NextDiT), SD VAE decoder."""
from . import flux, unet, wan, vae, zimage  # noqa: F401

FAMILIES = {
    "flux": flux.Flux,
    "unet": unet.UNetModel,
    "wan": wan.WanModel,
    "zimage": zimage.ZImageModel,
    "vae": vae.VAEDecoder,
}
