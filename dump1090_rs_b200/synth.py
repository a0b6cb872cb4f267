"""Synthetic 2.4 Msps CS16 IQ for BASELINE.json configs 3-5.

Noise: an RTL-SDR-like 8-bit front end, u8 = clip(round(127.4 + N(0, sigma)), 0, 255),
cs16 = trunc((u8 - 127.4) / 128 * 32767) -- the mapping that reproduces the value set
of the reference's test_iq captures.  Signals: DF17 extended squitters with a valid
CRC-24, built as a 12 MHz envelope (0.5 us = 6 ticks), delayed by a sub-sample phase,
box-averaged by 5 down to 2.4 Msps and added to I/Q before quantisation.

Data generation only: nothing here demodulates.
"""
from __future__ import annotations

import numpy as np

SAMPLES = 131072


def _crc24(data: bytes) -> int:
    """Mode-S parity of `data` (the 11 data bytes of a DF17): remainder mod 0x1FFF409."""
    rem = 0
    for b in data:
        rem ^= b << 16
        for _ in range(8):
            rem = ((rem << 1) ^ 0xFFF409) & 0xFFFFFF if rem & 0x800000 else (rem << 1) & 0xFFFFFF
    return rem


def df17_message(icao: int, me: bytes, ca: int = 5) -> bytes:
    body = bytes([(17 << 3) | ca, (icao >> 16) & 0xFF, (icao >> 8) & 0xFF, icao & 0xFF]) + bytes(me[:7])
    p = _crc24(body)
    return body + bytes([(p >> 16) & 0xFF, (p >> 8) & 0xFF, p & 0xFF])


def envelope_12mhz(msg: bytes) -> np.ndarray:
    """On/off envelope of preamble + PPM bits at 12 MHz (1 us = 12 ticks)."""
    nbits = 8 * len(msg)
    env = np.zeros(12 * (8 + nbits) + 24, dtype=np.float64)
    for start_us in (0.0, 1.0, 3.5, 4.5):
        s = int(round(start_us * 12))
        env[s:s + 6] = 1.0
    for n in range(nbits):
        bit = (msg[n >> 3] >> (7 - (n & 7))) & 1
        s = 12 * (8 + n) + (0 if bit else 6)
        env[s:s + 6] = 1.0
    return env


def waveform_2400(msg: bytes, phase_ticks: int) -> np.ndarray:
    env = envelope_12mhz(msg)
    d = np.concatenate([np.zeros(phase_ticks), env, np.zeros(10)])
    n = d.size // 5
    return d[: 5 * n].reshape(n, 5).mean(axis=1)


def quantise(fi: np.ndarray, fq: np.ndarray) -> np.ndarray:
    """float 'u8 domain' I/Q -> int16 (re, im) pairs in memory order."""
    out = np.empty((fi.size, 2), dtype=np.int16)
    for col, f in ((0, fi), (1, fq)):
        u8 = np.clip(np.rint(f), 0, 255)
        out[:, col] = np.trunc((u8 - 127.4) / 128.0 * 32767.0).astype(np.int16)
    return out


def noise_buffer(rng: np.random.Generator, n: int = SAMPLES, sigma: float = 5.5):
    return 127.4 + rng.normal(0.0, sigma, n), 127.4 + rng.normal(0.0, sigma, n)


def make_buffer(seed: int, index: int, n: int = SAMPLES, msgs_per_buffer: int = 0, sigma: float = 5.5,
                icao_pool: int = 64, amplitude: float = 0.25):
    """One buffer: seed 1090-style base seed, stream = buffer index.  Returns
    (int16 [n,2] (re,im), list of injected (offset, msg bytes))."""
    rng = np.random.default_rng([seed, index])
    fi, fq = noise_buffer(rng, n, sigma)
    injected = []
    if msgs_per_buffer:
        grid = 400
        slots = rng.permutation(max((n - 400) // grid, 1))[:msgs_per_buffer]
        for s in sorted(int(x) for x in slots):
            icao = 0xA00000 + int(rng.integers(0, icao_pool)) * 0x111
            msg = df17_message(icao, bytes(rng.integers(0, 256, 7, dtype=np.uint8)))
            wf = waveform_2400(msg, int(rng.integers(0, 5)))
            off = s * grid + int(rng.integers(0, 60))
            m = min(wf.size, n - off)
            th = rng.uniform(0, 2 * np.pi)
            fi[off:off + m] += amplitude * 128.0 * wf[:m] * np.cos(th)
            fq[off:off + m] += amplitude * 128.0 * wf[:m] * np.sin(th)
            injected.append((off, msg))
    return quantise(fi, fq), injected


def make_batch(seed: int, n_buffers: int, n: int = SAMPLES, msgs_per_buffer: int = 0, first_index: int = 0,
               **kw) -> np.ndarray:
    out = np.empty((n_buffers, n, 2), dtype=np.int16)
    for b in range(n_buffers):
        out[b], _ = make_buffer(seed, first_index + b, n, msgs_per_buffer, **kw)
    return out


def full_range_buffer(seed: int, index: int, n: int = SAMPLES, sigma: float = 1400.0) -> np.ndarray:
    rng = np.random.default_rng([seed, index, 7])
    return np.clip(np.rint(rng.normal(0, sigma, (n, 2))), -32768, 32767).astype(np.int16)


def noise_batch_torch(seed: int, n_buffers: int, n: int = SAMPLES, device="cuda", sigma: float = 5.5,
                      chunk: int = 64):
    """The same noise model generated on the device with torch (bench-scale batches).
    Not bit-identical to make_batch (different RNG); parity checks pull buffers back
    and run the oracle on exactly these values."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_buffers, n, 2), dtype=torch.int16, device=device)
    for b0 in range(0, n_buffers, chunk):
        nb = min(chunk, n_buffers - b0)
        f = torch.randn((nb, n, 2), generator=g, device=device, dtype=torch.float32) * sigma + 127.4
        u8 = torch.clamp(torch.round(f), 0, 255)
        out[b0:b0 + nb] = torch.trunc((u8 - 127.4) / 128.0 * 32767.0).to(torch.int16)
    return out
