"""Host-side filter design (float64 / extended precision), no signal data touched here.

The reference designs its Butterworth filters as rounded ``(b, a)`` polynomials and
runs an 8th-order direct form (ref: preprocess/signal/frequency_filter.py:218-229).
A direct form is unusable in a chunk-parallel GPU scan (and already loses 1e-5 in
float64 for a 4 Hz notch), so the device runs a biquad
cascade realising the SAME rational function: the reference's own rounded
``(b, a)`` are factored here with mpmath at 80 digits.  Everything the kernel
needs besides the signal is produced here:

* ``sos``   biquad coefficients, rounded once from the extended-precision roots;
* ``zi``    scipy's ``lfilter_zi(b, a)`` (with its round-off, which the reference
            output contains) mapped into cascade state coordinates through the
            observability matrices of the two realisations;
* ``M``     the cascade state-transition matrix to the power ``chunk``;
* ``tail``  how many trailing samples of a chunk still influence its end state.
"""
from __future__ import annotations

import functools
import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import mpmath as mp
import numpy as np
from scipy import signal as sp_signal

_DPS = 80
NUM_SMS = 148
SOS_THREADS = 512
SUB = 16


# ----------------------------------------------------------------------------- SOS
def _pair_roots(roots):
    """Group roots into conjugate pairs / real pairs; returns list of (r1, r2) (r2 may be None)."""
    roots = list(roots)
    tol = mp.mpf(10) ** (-_DPS // 2)
    cplx = [r for r in roots if abs(mp.im(r)) > tol * max(1, abs(r))]
    real = sorted((mp.re(r) for r in roots if abs(mp.im(r)) <= tol * max(1, abs(r))))
    pairs = []
    upper = sorted((r for r in cplx if mp.im(r) > 0), key=lambda r: (mp.re(r), mp.im(r)))
    lower = [r for r in cplx if mp.im(r) < 0]
    if len(upper) != len(lower):
        raise ValueError("complex roots do not come in conjugate pairs")
    for r in upper:
        j = min(range(len(lower)), key=lambda i: abs(lower[i] - mp.conj(r)))
        pairs.append((r, lower.pop(j)))
    while len(real) >= 2:
        # pair the extremes so a double zero at +1 and -1 becomes (1 - z^-2)
        pairs.append((real.pop(0), real.pop(-1)))
    if real:
        pairs.append((real.pop(), None))
    return pairs


def _roots(coeffs):
    """Roots of a real polynomial: mpmath at high precision; exact multiple roots (e.g. the
    (1 + z^-1)^4 numerator of a low-pass) make Durand-Kerner crawl, and there a backward-stable
    float64 companion-matrix solve is all that is needed (the product of the factors
    reproduces the coefficients to ~1e-15, which is what the transfer function sees)."""
    if len(coeffs) <= 1:
        return []
    try:
        return mp.polyroots(coeffs, maxsteps=400, extraprec=4 * _DPS)
    except mp.libmp.libhyper.NoConvergence:
        r = np.roots(np.array([float(c) for c in coeffs], dtype=np.float64))
        return [mp.mpc(float(v.real), float(v.imag)) for v in r]


def _quad(pair):
    r1, r2 = pair
    if r2 is None:
        return [mp.mpf(1), -mp.re(r1), mp.mpf(0)]
    return [mp.mpf(1), -mp.re(r1 + r2), mp.re(r1 * r2)]


def ba_to_sos(b: Sequence[float], a: Sequence[float]) -> np.ndarray:
    """Factor rounded ``(b, a)`` into biquads with extended-precision root finding."""
    with mp.workdps(_DPS):
        b = [mp.mpf(float(v)) for v in np.atleast_1d(b)]
        a = [mp.mpf(float(v)) for v in np.atleast_1d(a)]
        while len(b) > 1 and b[-1] == 0 and len(b) > len(a):
            b.pop()
        n = max(len(a), len(b)) - 1
        gain = b[0] / a[0]
        lead = 0
        while b[lead] == 0:          # leading zeros = pure delay; not produced by butter
            lead += 1
        if lead:
            raise ValueError("numerator with leading zeros is not supported")
        zeros = _roots(b)
        poles = _roots(a)
        zp, pp = _pair_roots(zeros), _pair_roots(poles)
        # most resonant poles first, each with its nearest remaining zero pair
        pp.sort(key=lambda pr: -max(abs(pr[0]), abs(pr[1]) if pr[1] is not None else 0))
        nsec = max(len(zp), len(pp))
        sos = np.zeros((nsec, 6))
        for j in range(nsec):
            den = _quad(pp[j]) if j < len(pp) else [mp.mpf(1), mp.mpf(0), mp.mpf(0)]
            if zp:
                ref = pp[j][0] if j < len(pp) else mp.mpf(0)
                i = min(range(len(zp)), key=lambda i: abs(zp[i][0] - ref))
                num = _quad(zp.pop(i))
            else:
                num = [mp.mpf(1), mp.mpf(0), mp.mpf(0)]
            if j == 0:
                num = [gain * v for v in num]
            sos[j, :3] = [float(v) for v in num]
            sos[j, 3:] = [float(v) for v in den]
        return sos


# ------------------------------------------------------------------ state space
def _cascade_step(sos, state, u):
    """One sample through the DF2T cascade, mirroring csrc/sosfilt.cu::sos_step."""
    state = [list(s) for s in state]
    for j, (b0, b1, b2, a0, a1, a2) in enumerate(sos):
        y = b0 * u + state[j][0]
        state[j][0] = b1 * u + state[j][1] - a1 * y
        state[j][1] = b2 * u - a2 * y
        u = y
    return state, u


def cascade_state_space(sos):
    """(A, B, C, D) of the cascade with state order [s1_0, s2_0, s1_1, s2_1, ...] (mpmath)."""
    nsec = len(sos)
    ns = 2 * nsec
    S = [[mp.mpf(float(v)) for v in row] for row in sos]
    S = [[r[0] / r[3], r[1] / r[3], r[2] / r[3], mp.mpf(1), r[4] / r[3], r[5] / r[3]] for r in S]
    zero = [[mp.mpf(0), mp.mpf(0)] for _ in range(nsec)]
    A = mp.zeros(ns, ns)
    Cm = mp.zeros(1, ns)
    for i in range(ns):
        st = [list(s) for s in zero]
        st[i // 2][i % 2] = mp.mpf(1)
        st2, y = _cascade_step(S, st, mp.mpf(0))
        for k in range(ns):
            A[k, i] = st2[k // 2][k % 2]
        Cm[0, i] = y
    st2, D = _cascade_step(S, zero, mp.mpf(1))
    B = mp.matrix([st2[k // 2][k % 2] for k in range(ns)])
    return A, B, Cm, D


def _direct_state_space(b, a):
    """DF2T direct form as scipy's lfilter runs it (a[0] == 1 after normalisation)."""
    n = max(len(a), len(b)) - 1
    b = list(b) + [mp.mpf(0)] * (n + 1 - len(b))
    a = list(a) + [mp.mpf(0)] * (n + 1 - len(a))
    b = [v / a[0] for v in b]
    a = [v / a[0] for v in a]
    A = mp.zeros(n, n)
    for i in range(n):
        A[i, 0] = -a[i + 1]
        if i + 1 < n:
            A[i, i + 1] = 1
    Cm = mp.zeros(1, n)
    Cm[0, 0] = 1
    return A, Cm


def _observability(A, Cm, rows):
    O = mp.zeros(rows, A.rows)
    v = Cm.copy()
    for r in range(rows):
        for c in range(A.rows):
            O[r, c] = v[0, c]
        v = v * A
    return O


def zi_to_cascade(b, a, sos, zi_direct) -> np.ndarray:
    """Cascade state with the same zero-input response as the direct-form state ``zi_direct``."""
    with mp.workdps(_DPS):
        bm = [mp.mpf(float(v)) for v in np.atleast_1d(b)]
        am = [mp.mpf(float(v)) for v in np.atleast_1d(a)]
        Ad, Cd = _direct_state_space(bm, am)
        Ac, _, Cc, _ = cascade_state_space(sos)
        n = Ad.rows
        ns = Ac.rows
        # a first-order section never uses its second state: drop dead columns
        live = [i for i in range(ns)
                if not (i % 2 == 1 and sos[i // 2][2] == 0 and sos[i // 2][5] == 0)]
        if len(live) != n:
            raise ValueError(f"cascade has {len(live)} live states for a filter of order {n}")
        rows = n
        Od = _observability(Ad, Cd, rows)
        Oc_full = _observability(Ac, Cc, rows)
        Oc = mp.zeros(rows, n)
        for r in range(rows):
            for j, i in enumerate(live):
                Oc[r, j] = Oc_full[r, i]
        rhs = Od * mp.matrix([mp.mpf(float(v)) for v in zi_direct])
        sol = mp.lu_solve(Oc, rhs)
        out = np.zeros(ns)
        for j, i in enumerate(live):
            out[i] = float(sol[j])
        return out.reshape(-1, 2)


def chunk_matrix(sos, chunk: int) -> np.ndarray:
    """A^chunk of the cascade, by repeated squaring in extended precision."""
    with mp.workdps(50):
        A, _, _, _ = cascade_state_space(sos)
        R = mp.eye(A.rows)
        P = A.copy()
        e = int(chunk)
        while e:
            if e & 1:
                R = R * P
            P = P * P
            e >>= 1
        return np.array([[float(R[i, j]) for j in range(A.cols)] for i in range(A.rows)])


def response_tail(sos, limit: int, eps: float = 1e-18) -> int:
    """Smallest multiple of SUB samples after which the zero-input response is < eps (<= limit)."""
    with mp.workdps(30):
        A, _, _, _ = cascade_state_space(sos)
        An = np.array([[float(A[i, j]) for j in range(A.cols)] for i in range(A.rows)])
    P = np.linalg.matrix_power(An, SUB)
    Q = np.eye(An.shape[0])
    steps = 0
    while steps * SUB < limit:
        Q = Q @ P
        steps += 1
        if np.max(np.abs(Q)) < eps:
            break
    return min(max(steps * SUB, SUB), limit)


def choose_chunk(C: int, T: int, chunk: Optional[int] = None) -> int:
    """Chunk length so that C * ceil(T/chunk) chunk-threads fill one wave of 148 SMs x 2 CTAs x 512."""
    if chunk is not None:
        if chunk % SUB:
            raise ValueError(f"chunk must be a multiple of {SUB}")
        return int(chunk)
    items = NUM_SMS * 2 * SOS_THREADS
    n_chunks = max(1, items // max(C, 1))
    L = -(-T // n_chunks)
    L = max(-(-L // SUB) * SUB, 1024)
    return int(L)


WARM_EPS = 1e-10          # zero-input response bound that makes a zero-state warm-up exact enough
WARM_MAX_OVERHEAD = 0.75  # use the warm-up path only while tail / chunk stays below this


def choose_warm_chunk(C: int, T: int, threads_per_sm: int = 512) -> int:
    """Chunk length of the warm-up path: C * ceil(T/chunk) chunk-threads = one wave of
    148 SMs x threads_per_sm.  Few warps per scheduler already saturate the FP64 pipe, and
    long chunks keep the redundant warm-up (tail / chunk) small."""
    n_chunks = max(1, (NUM_SMS * threads_per_sm) // max(C, 1))
    L = -(-T // n_chunks)
    return int(max(-(-L // SUB) * SUB, 1024))


def warm_tail(design: "SosDesign", limit: int) -> int:
    """Warm-up length (multiple of SUB) after which max|A^n| < WARM_EPS, or -1 if > limit."""
    t = _warm_tail(design.sos.tobytes(), design.nsec, int(limit))
    return t


@functools.lru_cache(maxsize=256)
def _warm_tail(sos_bytes: bytes, nsec: int, limit: int) -> int:
    sos = np.frombuffer(sos_bytes, dtype=np.float64).reshape(nsec, 6)
    t = response_tail(sos, limit + SUB, WARM_EPS)
    return t if t <= limit else -1


@dataclass(frozen=True)
class SosDesign:
    sos: np.ndarray            # (nsec, 6)
    zi: Optional[np.ndarray]   # (nsec, 2) cascade coordinates, None if causal
    padlen: int
    zero_phase: bool
    b: Optional[np.ndarray] = None
    a: Optional[np.ndarray] = None

    @property
    def nsec(self) -> int:
        return int(self.sos.shape[0])


NUMERATOR_TOL = 1e-6      # max-norm relative output difference (white noise) that still counts as "the same filter"


def numerator_deviation(sos_exact: np.ndarray, sos_alt: np.ndarray, n: int = 1 << 17) -> float:
    """max|y_exact - y_alt| / max|y_exact| for a fixed white-noise input through the two cascades
    (same poles, different numerator factorisations), float64, zero state."""
    u = np.random.default_rng(20230101).standard_normal(n)
    y1 = sp_signal.sosfilt(sos_exact, u)
    y2 = sp_signal.sosfilt(sos_alt, u)
    return float(np.max(np.abs(y1 - y2)) / np.max(np.abs(y1)))


@functools.lru_cache(maxsize=64)
def _butter_design(order: int, freqs: Tuple[float, ...], fs: float, btype: str, causal: bool) -> SosDesign:
    wn = np.asarray(freqs, dtype=float) / (0.5 * fs)
    wn = wn if wn.size > 1 else float(wn[0])
    if causal:
        # ref: frequency_filter.py:222-224 -- design-time SOS, zero initial state
        sos = sp_signal.butter(order, wn, btype=btype, output="sos")
        return SosDesign(np.ascontiguousarray(sos, dtype=np.float64), None, 0, False)
    # ref: frequency_filter.py:226-227 -- (b, a) + filtfilt defaults
    b, a = sp_signal.butter(order, wn, btype=btype)
    sos = ba_to_sos(b, a)
    # Cheaper numerator forms (csrc/sos_common.cuh).  Butterworth zeros sit ON the unit circle, so the
    # ideal numerator is b[0] (1 - z^-2)^order (band-pass) or b[0] (1 + beta z^-1 + z^-2)^nsec (band-stop,
    # low-/high-pass), and it equals scipy's ROUNDED b to ~1e-16 in every coefficient.  That is not the
    # same as "the same filter": next to a multiple zero the polynomial's VALUE is tiny (|B| ~ 1e-12 at the
    # edges of a 4-Hz notch at 3 kHz), so a 1e-16 change of the coefficients moves the response there by
    # 1e-4 -- the rounded b has its zeros on a ring of radius ~(1e-16)^(1/nsec) around the ideal one.
    # The reference's filter IS the rounded (b, a); its exact factors (`sos`, mpmath) reproduce the
    # long-double evaluation of the reference to 2e-7 (58-62 Hz at 3 kHz), the ideal form only to 5e-4.
    # So the ideal form is used only when the two give the same output to NUMERATOR_TOL on white noise.
    unit = None
    if btype == "bandpass" and sos.shape[0] == order:
        ideal = b[0] * np.poly(np.r_[np.ones(order), -np.ones(order)])
        if np.max(np.abs(ideal - b)) <= 4e-16 * np.max(np.abs(b)):
            unit = sos.copy()
            unit[:, :3] = [1.0, 0.0, -1.0]
            unit[0, :3] *= b[0]
    elif btype in ("bandstop", "lowpass", "highpass") and sos.shape[0] * 2 == len(b) - 1:
        nsec = sos.shape[0]
        beta = b[1] / (nsec * b[0])
        ideal = np.array([1.0])
        for _ in range(nsec):
            ideal = np.convolve(ideal, [1.0, beta, 1.0])
        if np.max(np.abs(b[0] * ideal - b)) <= 2e-15 * np.max(np.abs(b)):
            unit = sos.copy()
            unit[:, :3] = [1.0, beta, 1.0]
            unit[0, :3] *= b[0]
    if unit is not None and numerator_deviation(sos, unit) <= NUMERATOR_TOL:
        sos = unit
    zi_direct = sp_signal.lfilter_zi(b, a)          # the installed scipy's own arithmetic
    zi = zi_to_cascade(b, a, sos, zi_direct)
    padlen = 3 * max(len(a), len(b))
    return SosDesign(sos, zi, padlen, True, np.asarray(b), np.asarray(a))


def butter_design(freqs, fs, order=4, causal=False, filter_type="bandpass") -> SosDesign:
    freqs = tuple(float(f) for f in np.atleast_1d(np.asarray(freqs, dtype=float)))
    return _butter_design(int(order), freqs, float(fs), str(filter_type), bool(causal))


def pair_design(A: SosDesign, B: SosDesign):
    """Two 4-section zero-phase designs as ONE 8-section cascade for the fused sweep pair
    (csrc/sosfilt_pair.cu, split = 4): sections 0-3 = A, sections 4-7 = B; when B is of unit form its
    first section is made monic and its gain joins A's section 0.  Returns ``(design, tail_b)`` -- ``tail_b`` is the
    warm-up length of the second cascade (it forgets faster than the pair and joins the warm-up
    late) -- or None for a combination of numerator forms the kernel is not instantiated for."""
    return _pair_design(A.sos.tobytes(), B.sos.tobytes(), A.padlen, B.padlen)


@functools.lru_cache(maxsize=64)
def _pair_design(a_bytes: bytes, b_bytes: bytes, pad_a: int, pad_b: int):
    sa = np.frombuffer(a_bytes, dtype=np.float64).reshape(-1, 6).copy()
    sb = np.frombuffer(b_bytes, dtype=np.float64).reshape(-1, 6).copy()
    if sa.shape[0] != 4 or sb.shape[0] != 4:
        return None

    def form(sos):
        """Numerator form as the kernel classifies it (csrc/sos_common.cuh::unit_form)."""
        ok = sos[0, 0] != 0.0 and np.all(sos[1:, 0] == 1.0) and np.all(sos[:, 3] == 1.0)
        if ok and np.all(sos[:, 2] == sos[:, 0]):
            return 2
        if ok and np.all(sos[:, 2] == -sos[:, 0]) and np.all(sos[:, 1] == 0.0):
            return 5
        return 8 if ok else 0           # 8: monic general sections (exact factors), run in direct form II

    fa, fb = form(sa), form(sb)
    if (fa, fb) not in {(2, 5), (8, 5), (5, 2), (5, 8), (5, 5), (2, 2), (8, 8)}:
        return None
    if fb:                      # unit second half: monic, its gain joins section 0 of the first half
        gb = sb[0, 0]
        sb[0, :3] /= gb
        sa[0, :3] *= gb
    sos = np.ascontiguousarray(np.vstack([sa, sb]))
    zi = sp_signal.sosfilt_zi(sos)          # unit-step steady state in the kernel's DF2T coordinates
    dsg = SosDesign(sos, np.ascontiguousarray(zi), max(pad_a, pad_b), True)
    tb = _warm_tail(np.ascontiguousarray(np.frombuffer(b_bytes, dtype=np.float64)).tobytes(), 4, 1 << 30)
    return dsg, int(2 * tb)


# float32 second half of a cascade pair (csrc/sos_common.cuh: Bp32, ECOG_SOS_SPLIT_F32B): the delta-form
# recursion's round-off grows like eps32 / (1 - a2); measured 0.35 * 6e-8 / min(1 - a2) of the row maximum
# (3.5e-7 at 2 kHz, 4.5e-7 at 3 kHz for the 70-150 Hz band).  Below this bound the estimate stays <= ~1e-6.
BANDPASS_F32_MIN_E2 = 0.03


def bandpass_f32_ok(B: SosDesign) -> bool:
    """True when the 4-section band-pass ``B`` may run in float32 delta form inside a cascade pair: sections
    g (1 - z^-2) / (1 + a1 z^-1 + a2 z^-2) whose poles are far enough inside the unit circle."""
    sos = np.asarray(B.sos, dtype=np.float64)
    if sos.shape != (4, 6) or not np.all(sos[:, 3] == 1.0):
        return False
    if not (np.all(sos[:, 1] == 0.0) and np.all(sos[:, 2] == -sos[:, 0]) and np.all(sos[:, 0] != 0.0)):
        return False
    e2 = 1.0 - sos[:, 5]
    c1 = -(1.0 + sos[:, 4] + sos[:, 5])
    return bool(np.min(e2) >= BANDPASS_F32_MIN_E2 and np.all(c1 < 0.0))


@functools.lru_cache(maxsize=256)
def _chunk_ops(sos_bytes: bytes, nsec: int, chunk: int):
    sos = np.frombuffer(sos_bytes, dtype=np.float64).reshape(nsec, 6)
    return chunk_matrix(sos, chunk), response_tail(sos, chunk)


def chunk_ops(design: SosDesign, chunk: int):
    """(M = A^chunk, tail) for a design; cached."""
    return _chunk_ops(design.sos.tobytes(), design.nsec, int(chunk))


# ------------------------------------------------------------- Gaussian bank
def gaussian_bank(freq_ranges, f0=0.018, octspace=1.0 / 7.0,
                  filterbank_bias=math.log10(0.39), filterbank_slope=0.5):
    """Centre frequencies / widths of the Gaussian bank, ref: frequency_filter.py:121-152.
    Accepts the int spelling of example_config.yaml:36 that the reference rejects (B2)."""
    if isinstance(freq_ranges, tuple):
        freq_ranges = [freq_ranges]
    if isinstance(freq_ranges[0], (int, float, np.integer, np.floating)):
        freq_ranges = [tuple(freq_ranges)]
    cfs, sds = [], []
    for rng in freq_ranges:
        if len(rng) != 2:
            raise ValueError("Each frequency range must be a tuple of (min_freq, max_freq).")
        lo, hi = float(rng[0]), float(rng[1])
        max_oct = math.log2(hi / f0)
        f = f0
        while math.log2(f / f0) < max_oct:      # iterated product, like the reference
            if f >= lo:
                cfs.append(f)
                sds.append(10 ** (filterbank_bias + filterbank_slope * math.log10(f)))
            f = f * (2 ** octspace)
    return np.array(cfs), np.array(sds) * np.sqrt(2)
