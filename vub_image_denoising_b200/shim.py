"""Make the reference's import paths resolve to the B200 modules.

The reference scripts do ``from UNet.RDUNet_model import RDUNet``
(evaluate_Unet_diffusion/evaluate_model.py:18), ``from diffusion_denoising.diffusion_RDUnet import
RDUNet_T, DiffusionModel`` (evaluate_model.py:19, evaluate_SIDD/evaluate_SIDD.py:16,
evaluate_SIDD/benchmark.py:15) and ``from diffusion_denoising.Unet.Unet_model import RDUNet_T,
init_weights`` (diffusion_RDUnet.py:18).  ``install()`` registers light-weight stand-in modules under
those names in ``sys.modules`` so the scripts run unmodified on top of this package.
"""
from __future__ import annotations

import sys
import types

from . import diffusion, rdunet

__all__ = ["install", "uninstall", "ALIASES"]

ALIASES = ("UNet", "UNet.RDUNet_model", "diffusion_denoising", "diffusion_denoising.diffusion_RDUnet",
           "diffusion_denoising.Unet", "diffusion_denoising.Unet.Unet_model")


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__b200dn_shim__ = True
    return m


def install(force: bool = False, model=None) -> None:
    """Register the aliases. Existing non-shim modules are left alone unless force=True.
    ``model``: optional module to bind as the global model of the one-argument ``sidd.my_srgb_denoiser(x)``
    (the reference keeps it in a script-level global, evaluate_SIDD/benchmark.py:23-32)."""
    import torch

    if model is not None:
        from . import sidd
        sidd.set_default_model(model)

    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    mods = {
        "UNet": _module("UNet", __path__=[]),
        "UNet.RDUNet_model": _module("UNet.RDUNet_model", RDUNet=rdunet.RDUNet, init_weights=rdunet.init_weights,
                                     device=device),
        "diffusion_denoising": _module("diffusion_denoising", __path__=[]),
        "diffusion_denoising.Unet": _module("diffusion_denoising.Unet", __path__=[]),
        "diffusion_denoising.Unet.Unet_model": _module("diffusion_denoising.Unet.Unet_model",
                                                       RDUNet_T=rdunet.RDUNet_T, init_weights=rdunet.init_weights),
        "diffusion_denoising.diffusion_RDUnet": _module("diffusion_denoising.diffusion_RDUnet",
                                                        RDUNet_T=rdunet.RDUNet_T, init_weights=rdunet.init_weights,
                                                        DiffusionModel=diffusion.DiffusionModel, device=device),
    }
    for name, mod in mods.items():
        cur = sys.modules.get(name)
        if cur is not None and not getattr(cur, "__b200dn_shim__", False) and not force:
            continue
        sys.modules[name] = mod
    sys.modules["UNet"].RDUNet_model = sys.modules["UNet.RDUNet_model"]
    dd = sys.modules["diffusion_denoising"]
    dd.diffusion_RDUnet = sys.modules["diffusion_denoising.diffusion_RDUnet"]
    dd.Unet = sys.modules["diffusion_denoising.Unet"]
    dd.Unet.Unet_model = sys.modules["diffusion_denoising.Unet.Unet_model"]


def uninstall() -> None:
    for name in ALIASES:
        if getattr(sys.modules.get(name), "__b200dn_shim__", False):
            del sys.modules[name]
