"""ORACLE (test infrastructure, not product): ctypes front-end of oracle/noise_oracle.c."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        L = C.CDLL(str(_build.build()))
        L.b200dn_oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200dn_oracle_normals.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32]
        L.b200dn_oracle_degrade.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                            C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200dn_oracle_norm_to_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def philox4x32_10(counter, key) -> np.ndarray:
    ctr = np.asarray(counter, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    _load().b200dn_oracle_philox4x32_10(ctr.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def normals(n: int, seed: int, stream_id: int = 0) -> np.ndarray:
    z = np.empty(n, dtype=np.float32)
    _load().b200dn_oracle_normals(z.ctypes.data, n, seed & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF)
    return z


def degrade(clean_u8: np.ndarray, sigma, seed: int, stream_id: int = 0):
    """clean_u8 [B,H,W,C] uint8 -> (noisy_u8 [B,H,W,C], noisy_norm [B,C,H,W] f32, clean_norm [B,C,H,W] f32)."""
    clean_u8 = np.ascontiguousarray(clean_u8, dtype=np.uint8)
    B, H, W, Cn = clean_u8.shape
    sig = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float32), (B,)))
    z = normals(clean_u8.size, seed, stream_id)
    noisy_u8 = np.empty_like(clean_u8)
    noisy = np.empty((B, Cn, H, W), dtype=np.float32)
    clean = np.empty((B, Cn, H, W), dtype=np.float32)
    _load().b200dn_oracle_degrade(clean_u8.ctypes.data, B, H, W, Cn, sig.ctypes.data, seed & 0xFFFFFFFFFFFFFFFF,
                                  stream_id & 0xFFFFFFFF, z.ctypes.data, noisy_u8.ctypes.data, noisy.ctypes.data,
                                  clean.ctypes.data)
    return noisy_u8, noisy, clean


def norm_to_u8(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.float32)
    B, Cn, H, W = img.shape
    out = np.empty((B, H, W, Cn), dtype=np.uint8)
    _load().b200dn_oracle_norm_to_u8(img.ctypes.data, B, H, W, Cn, out.ctypes.data)
    return out
