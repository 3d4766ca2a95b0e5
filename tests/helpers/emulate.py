"""numpy models of the device algorithms, driven by the SAME host-built plans/tables the
kernels consume.  They let the CPU test-suite validate plan logic (chunk carries, FFT
decompositions, table layouts) without a GPU; the GPU tests then only have to prove the
kernels implement these simple semantics."""
from __future__ import annotations

import numpy as np
from scipy import signal as sp_signal


def sos_chunk_scan(x32: np.ndarray, design, chunk: int, M: np.ndarray, tail: int) -> np.ndarray:
    """Model of csrc/sosfilt.cu for one sweep pair.  x32: (C, T) float32."""
    sos = design.sos
    nsec = sos.shape[0]
    C, T = x32.shape
    P = design.padlen

    def sweep(u, s0):
        """u: (T,) float64 main samples; s0: (nsec,2) state at main start -> (y float32, end state)"""
        n_chunks = -(-T // chunk)
        g = [None] * n_chunks
        for k in range(n_chunks - 1):                      # tail pass, zero state
            seg = u[k * chunk:(k + 1) * chunk][chunk - tail:]
            _, zf = sp_signal.sosfilt(sos, seg, zi=np.zeros((nsec, 2)))
            g[k] = zf
        S = [s0]
        for k in range(n_chunks - 1):                      # scan
            S.append((M @ S[k].reshape(-1) + g[k].reshape(-1)).reshape(nsec, 2))
        y = np.empty(T, dtype=np.float32)
        zf = s0
        for k in range(n_chunks):                          # main pass
            seg = u[k * chunk:(k + 1) * chunk]
            out, zf = sp_signal.sosfilt(sos, seg, zi=S[k])
            y[k * chunk:k * chunk + len(seg)] = out.astype(np.float32)
        return y, zf

    out = np.empty_like(x32)
    for c in range(C):
        x = x32[c]
        if not design.zero_phase:
            out[c], _ = sweep(x.astype(np.float64), np.zeros((nsec, 2)))
            continue
        left = (np.float32(2.0) * x[0] - x[P:0:-1]).astype(np.float32)
        right = (np.float32(2.0) * x[-1] - x[-2:-(P + 2):-1]).astype(np.float32)
        s = design.zi * np.float64(left[0])
        _, s = sp_signal.sosfilt(sos, left.astype(np.float64), zi=s)
        yf, s_end = sweep(x.astype(np.float64), s)
        ypad, _ = sp_signal.sosfilt(sos, right.astype(np.float64), zi=s_end)      # float64 pad
        s = design.zi * ypad[-1]
        _, s = sp_signal.sosfilt(sos, ypad[::-1], zi=s)
        yb, _ = sweep(yf[::-1].astype(np.float64), s)
        out[c] = yb[::-1]
    return out


def sos_warm_model(x32: np.ndarray, design, chunk: int, tail: int) -> np.ndarray:
    """Model of csrc/sosfilt.cu::sos_warm_kernel: every chunk starts from a zero state `tail`
    samples early, or from the exact filtfilt start-up when the row edge is within `tail`."""
    sos = design.sos
    nsec = sos.shape[0]
    C, T = x32.shape
    P = design.padlen

    def sweep(u, s0):
        """u: (T,) samples in sweep order, s0: exact state at u[0] -> float32 result, exact end state"""
        y = np.empty(T, dtype=np.float32)
        n_chunks = -(-T // chunk)
        for k in range(n_chunks):
            a, b = k * chunk, min((k + 1) * chunk, T)
            if a <= tail:
                out, _ = sp_signal.sosfilt(sos, u[:b], zi=s0)
                y[a:b] = out[a:].astype(np.float32)
            else:
                out, _ = sp_signal.sosfilt(sos, u[a - tail:b], zi=np.zeros((nsec, 2)))
                y[a:b] = out[tail:].astype(np.float32)
        _, zf = sp_signal.sosfilt(sos, u, zi=s0)      # the last chunk's thread carries the true state on
        return y, zf

    out = np.empty_like(x32)
    for c in range(C):
        x = x32[c]
        if not design.zero_phase:
            out[c], _ = sweep(x.astype(np.float64), np.zeros((nsec, 2)))
            continue
        left = (np.float32(2.0) * x[0] - x[P:0:-1]).astype(np.float32)
        right = (np.float32(2.0) * x[-1] - x[-2:-(P + 2):-1]).astype(np.float32)
        s = design.zi * np.float64(left[0])
        _, s = sp_signal.sosfilt(sos, left.astype(np.float64), zi=s)
        yf, s_end = sweep(x.astype(np.float64), s)
        ypad, _ = sp_signal.sosfilt(sos, right.astype(np.float64), zi=s_end)
        s = design.zi * ypad[-1]
        _, s = sp_signal.sosfilt(sos, ypad[::-1], zi=s)
        yb, _ = sweep(yf[::-1].astype(np.float64), s)
        out[c] = yb[::-1]
    return out


def pair_ws_model(x32: np.ndarray, A, B, chunk: int, tail: int, tail_b: int) -> np.ndarray:
    """Model of csrc/sosfilt_pairws.cu for one row (float32 in, float32 out): ONE forward and ONE backward sweep of
    notch (float64 sections of design A) -> float32 -> band-pass (design B in float32 delta form, the product of the
    gains applied in float32 between the halves); every chunk starts from a ZERO state `tail` samples early (the
    band-pass `tail_b` samples early), at the row ends from a zero state at the edge.  The caller of the kernel
    overwrites the first / last `tail` samples of a row; the model leaves them as the kernel computes them."""
    T = x32.shape[0]
    sa = np.array(A.sos, dtype=np.float64)
    sb = np.array(B.sos, dtype=np.float64)
    gain = np.float32(np.prod(sa[:, 0]) * np.prod(sb[:, 0]))
    sa_monic = sa.copy()
    sa_monic[:, :3] /= sa[:, :1]
    c1 = (-(1.0 + sb[:, 4] + sb[:, 5])).astype(np.float32)
    e2 = (1.0 - sb[:, 5]).astype(np.float32)

    def bandpass32(v):
        w1 = np.zeros(4, dtype=np.float32)
        d = np.zeros(4, dtype=np.float32)
        y = np.empty(v.size, dtype=np.float32)
        for n in range(v.size):
            u = np.float32(v[n]) * gain
            for j in range(4):
                dn = np.float32(c1[j] * w1[j] + (u - e2[j] * d[j])) + d[j]
                u = dn + d[j]
                w1[j] = w1[j] + dn
                d[j] = dn
            y[n] = u
        return y

    def sweep(u):
        y = np.empty(T, dtype=np.float32)
        for a in range(0, T, chunk):
            b = min(a + chunk, T)
            lo = max(a - tail, 0)
            notch = sp_signal.sosfilt(sa_monic, u[lo:b].astype(np.float64)).astype(np.float32)      # zero state at lo
            lb = max(a - tail_b, lo)
            y[a:b] = bandpass32(notch[lb - lo:])[a - lb:]
        return y

    return sweep(sweep(x32)[::-1])[::-1]


# ----------------------------------------------------------------------------- FFT models
def _c(t):  # (n, 2) float32 table -> complex128
    return t[:, 0].astype(np.float64) + 1j * t[:, 1].astype(np.float64)


def tile_fft(cols: np.ndarray, axis_plan) -> np.ndarray:
    """Model of the in-shared-memory FFT of csrc/fft.cu: cols is (n, W) complex; rows are
    placed at perm[i], then in-place DIT stages with twiddles read from the W_n table."""
    n = axis_plan.n
    tw = _c(axis_plan.tw)
    buf = np.empty_like(cols, dtype=np.complex128)
    buf[axis_plan.perm] = cols
    L_prev = 1
    for r in axis_plan.radices:
        L = L_prev * r
        stride = n // L
        new = np.empty_like(buf)
        for g in range(n // L):
            for j in range(L_prev):
                base = g * L + j
                u = np.stack([buf[base + q * L_prev] * tw[(j * q) * stride] for q in range(r)])
                for p in range(r):
                    acc = 0
                    for q in range(r):
                        acc = acc + u[q] * np.exp(-2j * np.pi * p * q / r)
                    new[base + p * L_prev] = acc
        buf = new
        L_prev = L
    return buf


def big_fft(z: np.ndarray, plan) -> np.ndarray:
    """Model of the two-pass four-step forward FFT (natural order in and out)."""
    N, na, nb = plan.N, plan.a.n, plan.b.n
    hi, lo = _c(plan.tw_hi), _c(plan.tw_lo)
    S = lo.shape[0]
    x = z.reshape(na, nb)                                   # j = i * nb + c
    Y = tile_fft(x, plan.a)                                 # over i -> q, for every column c
    q = np.arange(na)[:, None]
    c = np.arange(nb)[None, :]
    e = q * c
    Y = Y * hi[e // S] * lo[e % S]
    inter = np.ascontiguousarray(Y.T)                       # transposed store: [c][q]
    if nb == 1:
        return inter.reshape(-1)
    X = tile_fft(inter, plan.b)                             # over c -> p, for every column q
    return X.reshape(-1)                                    # k = p * na + q


def halfband_stage_model(x32: np.ndarray, st: np.ndarray) -> np.ndarray:
    """One circular half-band stage of ecog_halfband2_decimate: y[n] = c x[2n] + sum_i g_i (x[2n-2i-1] + x[2n+2i+1])."""
    T = x32.shape[0]
    dt = x32.dtype                  # float32 like the kernel, or float64 as the exact sums
    n2 = 2 * np.arange(T // 2, dtype=np.int64)
    y = dt.type(st[0]) * x32[n2]
    for i, g in enumerate(st[1:]):
        y = y + dt.type(g) * x32[(n2 - 2 * i - 1) % T] + dt.type(g) * x32[(n2 + 2 * i + 1) % T]
    return y.astype(dt)


def fir_decimate_model(x32: np.ndarray, pre) -> np.ndarray:
    """Model of csrc/firdecim.cu for one row: circular FIR + decimate, float32 accumulation."""
    if pre.halfband is not None:
        return halfband_stage_model(halfband_stage_model(x32, pre.halfband[0]), pre.halfband[1])
    T = x32.shape[0]
    T1 = T // pre.D
    y = np.zeros(T1, dtype=np.float32)
    base = np.arange(T1, dtype=np.int64) * pre.D - pre.offset
    for j, h in enumerate(pre.taps):
        if h != 0:
            y += np.float32(h) * x32[(base + j) % T]
    return y


def resample_model(x32: np.ndarray, plan, bin_gain=None) -> np.ndarray:
    """Model of ecog_fft_resample for one row (float32 in, float32 out)."""
    T, num = plan.T, plan.num
    N, Nh = T // 2, num // 2
    z = x32[0::2].astype(np.float64) + 1j * x32[1::2].astype(np.float64)
    Z = big_fft(z, plan.fwd)
    m = min(num, T)
    k = np.arange(Nh + 1)
    Zk = Z[k % N]
    Zm = np.conj(Z[(N - k) % N])
    X = 0.5 * (Zk + Zm) - 0.5j * _c(plan.tw_T) * (Zk - Zm)
    Y = X * (num / T)
    if bin_gain is not None:
        Y = Y * bin_gain.astype(np.float64)
    Y[k > m // 2] = 0
    if m % 2 == 0 and num != T:
        Y[m // 2] *= 2.0 if num < T else 0.5
    Y[0] = Y[0].real
    Y[Nh] = Y[Nh].real
    kk = np.arange(Nh)
    E = 0.5 * (Y[kk] + np.conj(Y[Nh - kk]))
    O = 0.5 * (Y[kk] - np.conj(Y[Nh - kk])) * _c(plan.tw_num)
    G = E + 1j * O
    g = np.conj(big_fft(np.conj(G), plan.inv)) / Nh
    y = np.empty(num, dtype=np.float32)
    y[0::2] = g.real
    y[1::2] = g.imag
    return y


def czt_resample_model(x32: np.ndarray, plan, bin_gain=None) -> np.ndarray:
    """Model of ops._czt_resample for one row: the device runs exactly these steps with
    ecog_cplx_modulate / ecog_fft_c2c (complex64 there, the plan's float32 tables here)."""
    c = lambda t: t[:, 0].astype(np.float64) + 1j * t[:, 1].astype(np.float64)
    A = np.zeros(plan.M1, dtype=np.complex128)
    A[:plan.T] = x32.astype(np.float64) * c(plan.pre)
    c1 = np.fft.ifft(np.fft.fft(A) * c(plan.FB1)) * plan.M1
    mid = plan.mid if bin_gain is None else plan.mid * bin_gain[:plan.K].astype(np.float64)
    mid = mid.astype(np.complex64).astype(np.complex128)
    A2 = np.zeros(plan.M2, dtype=np.complex128)
    A2[:plan.K] = c1[:plan.K] * mid
    c2 = np.fft.ifft(np.fft.fft(A2) * c(plan.FB2)) * plan.M2
    return (c2[:plan.num] * c(plan.post)).real.astype(np.float32)


def fft4096_model(z: np.ndarray, tw: np.ndarray) -> np.ndarray:
    """Model of the 16x16x16 register FFT of csrc/hilbert.cu using the table from
    ecog_hilbert_twiddles (first two (16,256,2) float32 sections)."""
    tw = tw[: 2 * 16 * 256 * 2].reshape(2, 16, 256, 2).astype(np.float64)
    tw1 = tw[0, :, :, 0] + 1j * tw[0, :, :, 1]
    tw2 = tw[1, :, :, 0] + 1j * tw[1, :, :, 1]
    tid = np.arange(256)
    buf = np.zeros(4096, dtype=np.complex128)
    v = z.reshape(16, 256)                                  # v[n2, tid] = z[256 n2 + tid]
    A = np.fft.fft(v, axis=0) * tw1                         # k0, tid
    buf[(256 * np.arange(16)[:, None] + tid[None, :])] = A
    k0, n0 = tid >> 4, tid & 15
    idx = 256 * k0[None, :] + 16 * np.arange(16)[:, None] + n0[None, :]
    B = np.fft.fft(buf[idx], axis=0) * tw2                  # k1, tid
    buf[idx] = B
    idx3 = 16 * tid[None, :] + np.arange(16)[:, None]
    Xr = np.fft.fft(buf[idx3], axis=0)                      # k2, tid ; tid = 16 k0 + k1
    out = np.zeros(4096, dtype=np.complex128)
    k0, k1 = tid >> 4, tid & 15
    out[k0[None, :] + 16 * k1[None, :] + 256 * np.arange(16)[:, None]] = Xr
    return out


def fft4096_sparse_model(y256: np.ndarray, tw: np.ndarray) -> np.ndarray:
    """Model of the fast-path inverse of csrc/hilbert.cu (hilbert_env8_kernel): forward FFT of a
    4096-point signal that is non-zero only in its first 256 samples.  Pass 1 is the identity;
    pass 2 reads SG[16 n1 + n0] x twA[n1, tid], then x twB[k1, n0]; pass 3 as in the dense FFT."""
    tw = tw.astype(np.float64)
    sec = 16 * 256 * 2
    twA = tw[2 * sec:3 * sec].reshape(16, 256, 2)
    twA = twA[..., 0] + 1j * twA[..., 1]                    # [n1, tid]
    twB = tw[3 * sec:3 * sec + 512].reshape(16, 16, 2)
    twB = twB[..., 0] + 1j * twB[..., 1]                    # [k1, n0]
    tid = np.arange(256)
    k0, n0 = tid >> 4, tid & 15
    vin = y256[16 * np.arange(16)[:, None] + n0[None, :]] * twA          # [n1, tid]
    B = np.fft.fft(vin, axis=0) * twB[:, n0]                             # [k1, tid]
    buf = np.zeros(4096, dtype=np.complex128)
    buf[256 * k0[None, :] + 16 * np.arange(16)[:, None] + n0[None, :]] = B
    Xr = np.fft.fft(buf[16 * tid[None, :] + np.arange(16)[:, None]], axis=0)   # [k2, tid], tid = 16 k0 + k1
    out = np.zeros(4096, dtype=np.complex128)
    out[(tid >> 4)[None, :] + 16 * (tid & 15)[None, :] + 256 * np.arange(16)[:, None]] = Xr
    return out


def hilbert_block_model(x32: np.ndarray, plan, halo: int, envelope: bool = True) -> np.ndarray:
    """Model of ecog_hilbert_env for one row: overlap-save blocks with circular halo,
    two blocks per complex FFT, conj-forward inverse, mean folded into the gain."""
    N = 4096
    T = x32.shape[0]
    U = N - 2 * halo
    nblk = -(-T // U)
    y = np.zeros(T, dtype=np.float32)
    gain, shift, rows = plan
    g = gain.astype(np.float64)
    for b0 in range(0, nblk, 2):
        i = np.arange(N)
        a = x32[(b0 * U - halo + i) % T].astype(np.float64)
        has1 = b0 + 1 < nblk
        bb = x32[((b0 + 1) * U - halo + i) % T].astype(np.float64) if has1 else np.zeros(N)
        Z = np.fft.fft(a + 1j * bb)
        k = np.arange(1, N // 2)
        S0 = np.zeros(N // 2, dtype=np.complex128)
        S1 = np.zeros(N // 2, dtype=np.complex128)
        S0[k] = Z[k] + np.conj(Z[N - k])
        S1[k] = -1j * (Z[k] - np.conj(Z[N - k]))
        acc = [np.zeros(N), np.zeros(N)]
        for band in range(g.shape[0]):
            for sel, S in enumerate((S0, S1)):
                Yc = np.zeros(N, dtype=np.complex128)          # band shifted down to bin 0
                Yc[:rows * 256] = np.conj(S[shift[band]:shift[band] + rows * 256]) * g[band]
                zc = np.fft.fft(Yc)
                acc[sel] += np.abs(zc) if envelope else zc.real
        for sel in range(2 if has1 else 1):
            t = (b0 + sel) * U + np.arange(U)
            ok = t < T
            y[t[ok]] = acc[sel][halo + np.arange(U)][ok].astype(np.float32)
    return y
