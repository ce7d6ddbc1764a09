"""NumPy twin of the device logits generator -- TEST INFRASTRUCTURE (see reference_np.py).

The device generator (``als_synth_logits`` in csrc/synth.cu) and this function are
integer-only up to one exact int->float conversion and an exact power-of-two
scale, so they agree bit for bit on any platform.  Shapes follow the reference's
logits tensor (NHWC, ``active_learning.py:231``; producer ``enet_modules.py:1376-1380``).

Element (t, n, pixel, c) with global element index e = (n*P + pixel)*C + c:

    base  = sum of the 4 low bytes of mix(k_base ^ e*GOLD)   - 510      (approx normal)
    noise = sum of the 4 low bytes of mix(k_t    ^ e*GOLD)   - 510      (per MC sample)
    img   = mix(k_img ^ n*GOLD):  scale = 1 + (img & 7),  amp = (img >> 3) & 3
    bias  = low byte of mix(k_bias ^ (n*4096 + c)*GOLD)
    q     = 4*(scale*base + 4*amp*bias) + (T>1 ? scale*noise : 0)
    x     = q / 512                                                   (exact in fp32)

bf16 logits are the round-to-nearest-even truncation of x.
"""
from __future__ import annotations

import numpy as np

SEED_DEFAULT = 20191013
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_K_BASE = np.uint64(0x243F6A8885A308D3)
_K_IMG = np.uint64(0x13198A2E03707344)
_K_BIAS = np.uint64(0xA4093822299F31D0)
_K_T = np.uint64(0x082EFA98EC4E6C89)


def _mix(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(30))
        z = z * _M1
        z = z ^ (z >> np.uint64(27))
        z = z * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def _bytesum(h):
    m = np.uint64(0xFF)
    s = (h & m) + ((h >> np.uint64(8)) & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(24)) & m)
    return s.astype(np.int64) - 510


def synth_logits(T: int, n0: int, n_imgs: int, H: int, W: int, C: int,
                 seed: int = SEED_DEFAULT, dtype: str = "float32", squeeze_t: bool = True) -> np.ndarray:
    """Images n0 .. n0+n_imgs-1 of the synthetic pool as [T, n, H, W, C] (or
    [n, H, W, C] if T == 1 and squeeze_t).  dtype "float32" or "bfloat16"
    (bf16 is returned as uint16 bit patterns)."""
    P = H * W
    with np.errstate(over="ignore"):
        seed64 = np.uint64(seed)
        ks = _mix(seed64 * _GOLD + np.uint64(1))
        n = np.arange(n0, n0 + n_imgs, dtype=np.uint64)
        img = _mix((ks ^ _K_IMG) ^ (n * _GOLD))
        scale = (np.uint64(1) + (img & np.uint64(7))).astype(np.int64)            # [n]
        amp = ((img >> np.uint64(3)) & np.uint64(3)).astype(np.int64)             # [n]
        c = np.arange(C, dtype=np.uint64)
        bias = (_mix((ks ^ _K_BIAS) ^ ((n[:, None] * np.uint64(4096) + c[None, :]) * _GOLD))
                & np.uint64(0xFF)).astype(np.int64)                               # [n, C]
        e = ((n[:, None, None] * np.uint64(P) + np.arange(P, dtype=np.uint64)[None, :, None])
             * np.uint64(C) + c[None, None, :])                                   # [n, P, C]
        base = _bytesum(_mix((ks ^ _K_BASE) ^ (e * _GOLD)))
        q0 = 4 * (scale[:, None, None] * base + 4 * amp[:, None, None] * bias[:, None, :])
        out = np.empty((T, n_imgs, P, C), dtype=np.float32)
        for t in range(T):
            if T > 1:
                kt = _mix((ks ^ _K_T) + np.uint64(t) * _GOLD)
                noise = _bytesum(_mix(kt ^ (e * _GOLD)))
                q = q0 + scale[:, None, None] * noise
            else:
                q = q0
            out[t] = q.astype(np.float32) * np.float32(1.0 / 512.0)
    out = out.reshape(T, n_imgs, H, W, C)
    if dtype == "bfloat16":
        out = f32_to_bf16_bits(out)
    elif dtype != "float32":
        raise ValueError(dtype)
    if T == 1 and squeeze_t:
        out = out[0]
    return out


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16, returned as uint16 bit patterns."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = (b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)
    return r.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)
