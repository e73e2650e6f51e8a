"""Drop-in ``DiffusionModel`` — the deterministic cold-diffusion sampler around RDUNet_T.

Reference: diffusion_denoising/diffusion_RDUnet.py:27-55 (and the one-shot variant
diffusion_denoising/diffusion_RDUnet_direct.py:198-201).  Constructor, attributes (``unet``, mutable
``timesteps``) and methods (``forward_diffusion``, ``improved_sampling``, ``forward``) are unchanged.

B200 design of ``improved_sampling``:
  * both U-Net evaluations of a step use the same x_t (diffusion_RDUnet.py:44,47), so they run as ONE
    forward over a 2B batch whose image b reads x_t[b % B] and timestep plane t_all[step][b];
  * the 7 elementwise launches per step collapse into one ``b200dn_sampler_step`` kernel that follows
    the reference's fp32 operation order exactly;
  * the whole T-step loop (T x (1 + 68 + 1) launches) is captured once into a CUDA graph per
    (B, H, W, T, precision) and replayed, with no host round trip per step (the reference builds two
    ``torch.tensor([t/T], device=...)`` H2D copies per step, diffusion_RDUnet.py:43,46).
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .rdunet import RDUNet_T, _RDUNetBase

__all__ = ["DiffusionModel", "SAMPLER_PREC"]

# 40 chained forwards amplify rounding.  Measured on B200 against the fp32 CPU oracle, random-init
# RDUNet_T(32), T = 20, six seeds / two sizes (profiles/r01_precision_study.txt):
#   bf16    92.1 - 99.6 % of pixels within 1/255                      -> fails the 99.9 % bar
#   bf16x2  100 %, max err 2.7e-3 .. 5.4e-3, PSNR drift up to 0.021 dB  -> bf16 WEIGHT rounding alone is marginal
#   fp16    100 %, max err 1.8e-3 .. 2.8e-3, PSNR drift <= 0.0014 dB    -> default: same tcgen05 kind::f16 rate, 1 MMA
#   fp16x2  100 %, max err 4e-4 .. 7.5e-4 (fp16 hi+lo activations, 2 MMAs)
#   bf16x3  max err <= 5.4e-5 (fp32-validation build, 3 MMAs)
SAMPLER_PREC = os.environ.get("B200DN_SAMPLER_PREC", "fp16")
# fp16 storage has 5 exponent bits: activations beyond +-65504 saturate in the epilogue (cvt.rn.satfinite) instead of
# becoming inf.  Random-init and the reference's trained checkpoints keep O(1) activations, but nothing guarantees it
# for an arbitrary checkpoint, so every fp16 launch raises a device flag when it clamps a value and improved_sampling
# re-runs the call in this (bf16-range) mode when the flag is set.
SAT_FALLBACK_PREC = os.environ.get("B200DN_SAMPLER_FALLBACK_PREC", "bf16x2")


def _f32(v: float) -> float:
    """fp32 rounding of a python double — what PyTorch does with a python scalar in an fp32 tensor op."""
    return float(np.float32(v))


class _SamplerState:
    """Static buffers + captured graph of one (B, H, W, T, precision) sampling configuration."""

    def __init__(self, model: "DiffusionModel", B: int, H: int, W: int, T: int, precision: str, use_graph: bool):
        unet: _RDUNetBase = model.unet
        self.lib = _lib.lib()
        self.B, self.H, self.W, self.T = B, H, W, T
        dev = next(unet.parameters()).device
        self.device = dev
        self.plan = unet.plan(2 * B, H, W, precision)
        self.signature = self.plan.signature
        self.y = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        self.xa = torch.empty_like(self.y)
        self.xb = torch.empty_like(self.y)
        self.u = torch.empty((2 * B, 3, H, W), dtype=torch.float32, device=dev)
        # per step: B copies of fl32(t/T) then B copies of fl32((t-1)/T)  (diffusion_RDUnet.py:43,46).  Built with
        # device ops (an fp64 quotient rounded to fp32 is what torch.tensor([t / T]) holds) so that a first call made
        # inside a caller's graph capture needs no pageable H2D copy.
        self.coef = []
        for t in range(T, 0, -1):
            a_t, a_p = t / T, (t - 1) / T
            self.coef.append((_f32(1 - a_t), _f32(a_t), _f32(1 - a_p), _f32(a_p)))
        steps = torch.arange(T, 0, -1, dtype=torch.float64, device=dev)
        a_t32 = (steps / T).to(torch.float32).view(T, 1).expand(T, B)
        a_p32 = ((steps - 1) / T).to(torch.float32).view(T, 1).expand(T, B)
        self.t_all = torch.cat([a_t32, a_p32], dim=1).contiguous()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.result = self.y  # set by _enqueue
        self.state_id = ops.register_plan(self)     # handle the torch.ops.b200dn.improved_sampling custom op takes
        # a first call made while the caller is itself capturing a graph must not synchronise or start a nested
        # capture: it runs the launch list eagerly inside the caller's capture
        if use_graph and T > 0 and not torch.cuda.is_current_stream_capturing():
            # warm-up run on a side stream (lazy module loads, smem attribute), then capture
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                self._enqueue()
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g

    def _enqueue(self) -> None:
        lib, B = self.lib, self.B
        n = self.y.numel()
        cur, nxt = self.y, self.xa            # x_T = noisy image (never written: ownership stays with y)
        for i in range(self.T):
            self.plan.run(cur, self.u, x_batch=B, t_ptr=self.t_all[i].data_ptr(), t_strides=(1, 0, 0))
            c1, at, c2, ap = self.coef[i]
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = lib.b200dn_sampler_step(cur.data_ptr(), self.u.data_ptr(), self.u[B:].data_ptr(), self.y.data_ptr(),
                                         c1, at, c2, ap, nxt.data_ptr(), n, stream)
            _lib.check(rc, "sampler_step")
            cur, nxt = nxt, (self.xb if nxt is self.xa else self.xa)
        self.result = cur

    def sample(self, noisy: torch.Tensor) -> torch.Tensor:
        self.y.copy_(noisy)
        if self.plan.sat_flag is not None:
            self.plan.sat_flag.zero_()
        if self.graph is not None and not torch.cuda.is_current_stream_capturing():
            self.graph.replay()
        else:
            self._enqueue()
        return self.result.clone()

    def saturated(self) -> bool:
        """True if an fp16-stored activation was clamped at +-65504 (or was NaN) during the last ``sample``.
        Reads one int from the device (synchronises the current stream)."""
        return self.plan.sat_flag is not None and bool(int(self.plan.sat_flag.item()))


class DiffusionModel(nn.Module):
    def __init__(self, unet: nn.Module, timesteps: int = 20):
        super().__init__()
        self.unet = unet
        self.timesteps = timesteps
        self.precision = SAMPLER_PREC
        self.use_cuda_graph = os.environ.get("B200DN_GRAPH", "1") != "0"
        # fp16 saturation guard: check the device flag after every fp16 sampling call (one 4-byte D2H read) and re-run
        # at `saturation_fallback` precision if anything was clamped; `last_saturated` records what happened
        self.check_saturation = os.environ.get("B200DN_SAT_CHECK", "1") != "0"
        self.saturation_fallback = SAT_FALLBACK_PREC
        self.last_saturated = False
        self._states: dict = {}

    # caches must not survive .to()/.cuda() or pickling
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._states = {}
        return out

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_states"] = {}
        return state

    def forward_diffusion(self, clean_image, noisy_image, t):
        """alpha * noisy + (1 - alpha) * clean with alpha = t / timesteps (diffusion_RDUnet.py:33-36)."""
        alpha = t / self.timesteps
        if isinstance(alpha, torch.Tensor) or not (isinstance(clean_image, torch.Tensor) and clean_image.is_cuda
                                                   and clean_image.dtype == torch.float32
                                                   and clean_image.shape == noisy_image.shape):
            # per-sample tensor timesteps are the training path (diffusion_RDUnet.py:93): plain tensor algebra
            return alpha * noisy_image + (1 - alpha) * clean_image
        clean = clean_image.detach().contiguous()
        noisy = noisy_image.detach().to(torch.float32).contiguous()
        out = torch.empty_like(clean)
        with torch.cuda.device(clean.device):
            rc = _lib.lib().b200dn_lerp(clean.data_ptr(), noisy.data_ptr(), _f32(alpha), _f32(1 - alpha),
                                        out.data_ptr(), out.numel(),
                                        torch.cuda.current_stream(clean.device).cuda_stream)
        _lib.check(rc, "lerp")
        return out

    @torch.no_grad()
    def improved_sampling(self, noisy_image: torch.Tensor) -> torch.Tensor:
        """x_T = y; for t = T..1: x <- x - [(1-a_t) U(x,t) + a_t y] + [(1-a_{t-1}) U(x,t-1) + a_{t-1} y]
        (diffusion_RDUnet.py:38-50).  Returns a fresh tensor; the input is never written."""
        unet = self.unet
        if not isinstance(unet, RDUNet_T):
            return self._generic_sampling(noisy_image)
        y = unet._check_input(noisy_image)
        st = self._state(y, self.precision)
        out = torch.ops.b200dn.improved_sampling(y, st.state_id)
        self.last_saturated = False
        if (self.check_saturation and st.plan.sat_flag is not None and not torch.cuda.is_current_stream_capturing()
                and st.saturated()):
            self.last_saturated = True
            warnings.warn(f"improved_sampling: {self.precision} activations saturated at +-65504; re-running this "
                          f"call in {self.saturation_fallback} (set DiffusionModel.precision to avoid the retry)",
                          RuntimeWarning, stacklevel=2)
            st = self._state(y, self.saturation_fallback)
            out = torch.ops.b200dn.improved_sampling(y, st.state_id)
        return out

    def _state(self, y: torch.Tensor, precision: str) -> "_SamplerState":
        unet = self.unet
        B, _, H, W = y.shape
        T = int(self.timesteps)
        key = (B, H, W, T, precision, self.use_cuda_graph)
        with torch.cuda.device(y.device):
            st = self._states.get(key)
            if st is None or st.signature != unet._param_signature():
                if st is None and len(self._states) >= 2:
                    self._states.pop(next(iter(self._states)))
                st = _SamplerState(self, B, H, W, T, precision, self.use_cuda_graph)
                self._states[key] = st
        return st

    def _generic_sampling(self, noisy_image: torch.Tensor) -> torch.Tensor:
        """Any other ``unet(x, t)`` module: reference loop with the fused step kernel."""
        y = noisy_image.detach().to(torch.float32).contiguous()
        if not y.is_cuda:
            raise RuntimeError("vub_image_denoising_b200 runs on CUDA (sm_100) tensors only; there is no CPU fallback")
        lib = _lib.lib()
        T = int(self.timesteps)
        x = y
        with torch.cuda.device(y.device):
            for t in range(T, 0, -1):
                a_t, a_p = t / T, (t - 1) / T
                tt = torch.tensor([a_t], device=y.device).view(1, 1, 1, 1)
                tp = torch.tensor([a_p], device=y.device).view(1, 1, 1, 1)
                u1 = self.unet(x, tt).contiguous()
                u2 = self.unet(x, tp).contiguous()
                nxt = torch.empty_like(y)
                rc = lib.b200dn_sampler_step(x.data_ptr(), u1.data_ptr(), u2.data_ptr(), y.data_ptr(),
                                             _f32(1 - a_t), _f32(a_t), _f32(1 - a_p), _f32(a_p),
                                             nxt.data_ptr(), y.numel(),
                                             torch.cuda.current_stream(y.device).cuda_stream)
                _lib.check(rc, "sampler_step")
                x = nxt
        return x.clone() if x is y else x

    @torch.no_grad()
    def direct_sampling(self, noisy_image: torch.Tensor) -> torch.Tensor:
        """One-shot variant: unet(noisy, t=1) (diffusion_RDUnet_direct.py:198-201)."""
        t = torch.tensor([1.0], device=noisy_image.device).view(1, 1, 1, 1)
        return self.unet(noisy_image, t)

    def forward(self, clean_image, noisy_image, t):
        noisy_step_image = self.forward_diffusion(clean_image, noisy_image, t)
        return self.improved_sampling(noisy_step_image)
