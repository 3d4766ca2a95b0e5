"""Synthetic ECoG sessions: seeded per channel so any
channel subset of a large recording can be regenerated without the rest.

Host-side numpy; used by tests and by the CPU legs of ``bench.py``.  The GPU
bench fills its full-size input on the device (``device_session``).
"""
from __future__ import annotations

import numpy as np

SEED = 20230101
COMMON_IDX = 2 ** 31 - 1


def _pink(rng: np.random.Generator, T: int) -> np.ndarray:
    """Unit-RMS 1/f noise: white noise shaped by f^-1/2 in the rfft domain."""
    X = np.fft.rfft(rng.standard_normal(T))
    f = np.arange(X.size, dtype=np.float64)
    f[0] = 1.0
    y = np.fft.irfft(X / np.sqrt(f), n=T)
    return y / np.sqrt(np.mean(y * y))


def events(session: int, duration_s: float, n_events: int, n_tones=4, n_syll=2):
    """One-decimal onsets drawn without replacement from the 0.1 s grid, plus
    uniform tone (1..n_tones) and syllable (0..n_syll-1) labels."""
    rng = np.random.default_rng([SEED, session])
    grid = np.arange(300, int(10 * (duration_s - 5)))
    pick = np.sort(rng.choice(grid, n_events, replace=len(grid) < n_events))
    onsets = pick / 10.0
    tone = rng.integers(1, n_tones + 1, n_events)
    syll = rng.integers(0, n_syll, n_events)
    return onsets, tone, syll


def session_channels(session: int, channels, T: int, fs: float, n_channels_total: int,
                     onsets=None, tone=None, syll=None) -> np.ndarray:
    """float32 (len(channels), T) raw ECoG in microvolt scale."""
    t = np.arange(T) / fs
    common = 0.5 * 30.0 * _pink(np.random.default_rng([SEED, session, COMMON_IDX]), T)
    line2 = 3.0 * np.sin(2 * np.pi * 120 * t) + 1.0 * np.sin(2 * np.pi * 180 * t)
    eighth = max(n_channels_total // 8, 1)
    out = np.empty((len(channels), T), dtype=np.float32)
    for i, c in enumerate(channels):
        rng = np.random.default_rng([SEED, session, int(c)])
        x = 30.0 * _pink(rng, T) + common
        x += 10.0 * np.sin(2 * np.pi * 60 * t + rng.uniform(0, 2 * np.pi)) + line2
        if onsets is not None and c < 2 * eighth:
            amp_of = (lambda k: 1 + 0.5 * tone[k]) if c < eighth else (lambda k: 1 + 0.8 * syll[k])
            nb = int(0.3 * fs)
            win = np.hanning(nb)
            carrier = rng.standard_normal(T)
            Xc = np.fft.rfft(carrier)
            ff = np.fft.rfftfreq(T, 1 / fs)
            Xc[(ff < 70) | (ff > 150)] = 0
            hg = np.fft.irfft(Xc, n=T)
            hg *= 8.0 / np.sqrt(np.mean(hg * hg))
            for k, on in enumerate(onsets):
                s = int(round((on + 0.2) * fs))
                if s + nb <= T:
                    x[s:s + nb] += amp_of(k) * win * hg[s:s + nb]
        out[i] = x.astype(np.float32)
    return out


def session(session_idx: int, C: int, T: int, fs: float, n_events: int = 0):
    """Whole (C, T) session plus its event table (or None)."""
    ev = events(session_idx, T / fs, n_events) if n_events else (None, None, None)
    x = session_channels(session_idx, range(C), T, fs, C, *ev)
    return x, ev


def audio(session_idx: int, duration_s: float, sf: float = 24414.0625) -> np.ndarray:
    rng = np.random.default_rng([SEED, session_idx, 7])
    return rng.standard_normal((1, int(sf * duration_s))).astype(np.float32)


def device_session(C: int, T: int, fs: float, seed: int = 0, device="cuda", channel_seed=None):
    """Full-size synthetic input generated ON the device for the bench: white
    noise + shared common-mode + 60/120/180 Hz line noise, float32 (C, T).
    Generated in channel blocks to bound temporaries."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED + seed)
    x = torch.empty((C, T), dtype=torch.float32, device=device)
    t = torch.arange(T, device=device, dtype=torch.float64) / fs
    line = (10.0 * torch.sin(2 * np.pi * 60 * t) + 3.0 * torch.sin(2 * np.pi * 120 * t)
            + torch.sin(2 * np.pi * 180 * t)).to(torch.float32)
    del t
    common = 15.0 * torch.randn(T, generator=g, device=device, dtype=torch.float32)
    if channel_seed is not None:          # channel shard: shared terms above, private channel noise below
        g.manual_seed(SEED + 7919 * (int(channel_seed) + 1) + seed)
    blk = 16
    for c0 in range(0, C, blk):
        c1 = min(C, c0 + blk)
        x[c0:c1].normal_(0.0, 30.0, generator=g)
        x[c0:c1] += common
        x[c0:c1] += line
    return x
