"""Init-time basis / filter design (layer L2 of SURVEY.md §1).

Everything here runs once per module construction on the host, in float64
NumPy, and is cast to float32 the same way the reference casts, so that the
registered buffers are interchangeable with a reference nnAudio ``state_dict``.
The float64 *operation order* of every expression below follows the reference
on purpose: bit-identical fp32 buffers are part of the drop-in contract
 and IEEE arithmetic is order sensitive.

Reference behaviour restated here (``Installation/nnAudio/``):
  utils.py:241-393   create_fourier_kernels   -> :func:`fourier_basis`
  utils.py:399-473   create_cqt_kernels       -> :func:`cqt_bank`
  utils.py:476-495   get_window_dispatch      -> :func:`_window_dispatch`
  utils.py:562-596   create_lowpass_filter    -> :func:`lowpass_fir`
  utils.py:599-677   early-downsample helpers -> :func:`early_downsample_plan`
  librosa_functions.py:201-486  mel filterbank      -> :func:`mel_filterbank`
  librosa_functions.py:13-198   gammatone weights   -> :func:`gammatone_filterbank`
  librosa_functions.py:493-564  pad_center          -> :func:`center_pad`
  mel.py:281-307     FFT-route DCT-II (ortho) -> :func:`dct2_ortho_matrix` (dense)
"""
from __future__ import annotations

import math
import warnings

import numpy as np
from scipy import signal as _sps

__all__ = [
    "center_pad",
    "fourier_basis",
    "cqt_bank",
    "lowpass_fir",
    "early_downsample_plan",
    "mel_filterbank",
    "gammatone_filterbank",
    "dct2_ortho_matrix",
    "ParameterError",
]


class ParameterError(ValueError):
    """The reference raises an *undefined* ``ParameterError`` (a NameError in
    practice); we define it as a ValueError subclass."""


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def center_pad(vec: np.ndarray, size: int) -> np.ndarray:
    """Zero-pad a 1-D window to ``size`` with the data centred
    (librosa_functions.py:493-564; left pad = (size-n)//2)."""
    n = vec.shape[-1]
    left = int((size - n) // 2)
    if left < 0:
        raise ParameterError(
            "Target size ({:d}) must be at least input size ({:d})".format(size, n)
        )
    return np.pad(vec, (left, int(size - n - left)), mode="constant")


def next_pow2_exponent(a) -> int:
    """utils.py:128-148 — ceil(log2(a))."""
    return int(np.ceil(np.log2(a)))


def _window_dispatch(window, n: int, fftbins: bool = True):
    """utils.py:476-495. String -> scipy window; ('gaussian', att_dB) tuple ->
    gaussian with sigma derived from the attenuation.  Other tuples / floats
    fall through to ``None`` exactly like the reference (which then fails in
    the caller); unsupported types raise."""
    if isinstance(window, str):
        return _sps.get_window(window, n, fftbins=fftbins)
    if isinstance(window, tuple):
        if window[0] == "gaussian":
            assert window[1] >= 0
            sigma = np.floor(-n / 2 / np.sqrt(-2 * np.log(10 ** (-window[1] / 20))))
            return _sps.get_window(("gaussian", sigma), n, fftbins=fftbins)
        return None
    if isinstance(window, float):
        return None
    raise Exception(
        "The function get_window from scipy only supports strings, tuples and floats."
    )


# --------------------------------------------------------------------------- #
# Fourier basis  (STFT / Mel / MFCC / Gammatone)
# --------------------------------------------------------------------------- #
def fourier_basis(
    n_fft,
    win_length=None,
    freq_bins=None,
    fmin=50,
    fmax=6000,
    sr=44100,
    freq_scale="linear",
    window="hann",
    verbose=True,
):
    """Un-windowed sin/cos DFT rows plus the centred window.

    Returns ``(wsin, wcos, bins2freq, binslist, window_mask)`` with
    ``wsin/wcos`` float32 ``(freq_bins, 1, n_fft)`` and ``window_mask``
    float32 ``(n_fft,)``.  The caller multiplies basis by window **in fp32**
    (stft.py:230-232).

    Bin k of the basis is ``sin/cos(2*pi*b_k*s/n_fft)`` for ``s=0..n_fft-1``
    where ``b_k`` is the (possibly fractional) digital bin:
      'no'     b_k = k                                   (utils.py:379-384)
      'linear' b_k = k*scaling + start_bin               (utils.py:319-337)
      'log'    b_k = exp(k*scaling) * start_bin          (utils.py:339-357)
      'log2'   b_k = 2**(k*scaling) * start_bin          (utils.py:359-377)
    """
    if freq_bins is None:
        freq_bins = n_fft // 2 + 1
    if win_length is None:
        win_length = n_fft

    s = np.arange(0, n_fft, 1.0)
    win = _sps.get_window(window, int(win_length), fftbins=True)
    win = center_pad(win, n_fft)

    # The F digital-bin values are formed with scalar arithmetic (exactly the
    # scalar libm/numpy calls the reference makes per bin); only the big
    # (F x n_fft) sin/cos evaluation is vectorised.
    if freq_scale in ("linear", "log", "log2"):
        if verbose:
            print(
                f"sampling rate = {sr}. Please make sure the sampling rate is correct in order to"
                f"get a valid freq range"
            )
        start_bin = fmin * n_fft / sr
        if freq_scale == "linear":
            scaling = (fmax - fmin) * (n_fft / sr) / freq_bins
            digital = [k * scaling + start_bin for k in range(freq_bins)]
        elif freq_scale == "log":
            scaling = np.log(fmax / fmin) / freq_bins
            digital = [np.exp(k * scaling) * start_bin for k in range(freq_bins)]
        else:
            scaling = np.log2(fmax / fmin) / freq_bins
            digital = [2 ** (k * scaling) * start_bin for k in range(freq_bins)]
        bins2freq = [d * sr / n_fft for d in digital]
        binslist = list(digital)
        coef = np.array([2 * np.pi * d for d in digital], dtype=np.float64)
    elif freq_scale == "no":
        bins2freq = [k * sr / n_fft for k in range(freq_bins)]
        binslist = list(range(freq_bins))
        coef = np.array([2 * np.pi * k for k in range(freq_bins)], dtype=np.float64)
    else:
        # The reference only prints a hint and returns uninitialised memory
        # (utils.py:385-386); failing loudly is the one deliberate deviation.
        raise ValueError(
            "Please select the correct frequency scale: 'linear', 'log', 'log2' or 'no'"
        )

    arg = coef[:, None] * s[None, :] / n_fft
    wsin = np.sin(arg).astype(np.float32)[:, None, :]
    wcos = np.cos(arg).astype(np.float32)[:, None, :]
    return wsin, wcos, bins2freq, binslist, win.astype(np.float32)


# --------------------------------------------------------------------------- #
# CQT wavelet bank  (CQT1992v2 / CQT2010v2 / VQT)
# --------------------------------------------------------------------------- #
def cqt_bank(
    Q,
    fs,
    fmin,
    n_bins=84,
    bins_per_octave=12,
    norm=1,
    window="hann",
    fmax=None,
    topbin_check=True,
    gamma=0,
):
    """Time-domain constant-Q wavelets, centred in a power-of-two frame.

    Returns ``(bank complex64 (n_bins, width), width, lengths float64
    (n_bins,), freqs float64 (n_bins,))``.  The reference returns lengths as a
    float32 torch tensor (utils.py:473); the module wraps ours the same way.

    Bin k: ``f_k = fmin*2^(k/bpo)``, ``l_k = ceil(Q*fs/(f_k + gamma/alpha))``,
    wavelet ``w(l_k) * exp(j*2*pi*f_k*n/fs) / l_k`` on ``n in [-l_k//2, l_k//2)``,
    L1/L2-normalised, placed at ``start = ceil(width/2 - l_k/2)`` (minus one for
    odd ``l_k``) -- utils.py:443-469.
    """
    if (fmax is not None) and (n_bins is None):
        n_bins = np.ceil(bins_per_octave * np.log2(fmax / fmin))
    elif (fmax is None) and (n_bins is not None):
        pass
    else:
        warnings.warn("If fmax is given, n_bins will be ignored", SyntaxWarning)
        n_bins = np.ceil(bins_per_octave * np.log2(fmax / fmin))
    freqs = fmin * 2.0 ** (np.r_[0:n_bins] / np.double(bins_per_octave))

    if np.max(freqs) > fs / 2 and topbin_check:
        raise ValueError(
            "The top bin {}Hz has exceeded the Nyquist frequency, \
                          please reduce the n_bins".format(
                np.max(freqs)
            )
        )

    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    lengths = np.ceil(Q * fs / (freqs + gamma / alpha))
    width = int(2 ** (np.ceil(np.log2(int(max(lengths))))))

    bank = np.zeros((int(n_bins), width), dtype=np.complex64)
    for k in range(int(n_bins)):
        f_k = freqs[k]
        l_k = lengths[k]  # float64, integral valued
        start = int(np.ceil(width / 2.0 - l_k / 2.0))
        if l_k % 2 == 1:
            start -= 1
        taps = _window_dispatch(window, int(l_k), fftbins=True)
        n = np.r_[-l_k // 2 : l_k // 2]
        wavelet = taps * np.exp(n * 1j * 2 * np.pi * f_k / fs) / l_k
        if norm:
            bank[k, start : start + int(l_k)] = wavelet / np.linalg.norm(wavelet, norm)
        else:
            bank[k, start : start + int(l_k)] = wavelet
    return bank, width, lengths, freqs


def lowpass_fir(band_center=0.5, kernel_length=256, transition_bandwidth=0.03):
    """Anti-alias FIR for the octave pyramid (utils.py:562-596): firwin2 with
    unit gain up to ``bc/(1+tb)`` and zero gain from ``bc*(1+tb)``."""
    pass_max = band_center / (1 + transition_bandwidth)
    stop_min = band_center * (1 + transition_bandwidth)
    taps = _sps.firwin2(kernel_length, [0.0, pass_max, stop_min, 1.0], [1.0, 1.0, 0.0, 0.0])
    return taps.astype(np.float32)


def early_downsample_plan(sr, hop_length, fmax_t, Q, n_octaves):
    """utils.py:599-677 (librosa's early-downsample rule).

    Returns ``(new_sr, new_hop, factor, fir_or_None)``; ``factor == 1`` means
    inactive.  count = min(max(0, ceil(log2(0.85*nyq/cutoff)) - 2),
                           max(0, ceil(log2(hop)) - n_octaves + 1)).
    """
    window_bandwidth = 1.5  # hann
    cutoff = fmax_t * (1 + 0.5 * window_bandwidth / Q)
    nyquist = sr // 2
    c1 = max(0, int(np.ceil(np.log2(0.85 * nyquist / cutoff)) - 1) - 1)
    c2 = max(0, next_pow2_exponent(hop_length) - n_octaves + 1)
    factor = 2 ** min(c1, c2)
    new_hop = hop_length // factor
    new_sr = sr / float(factor)
    fir = None
    if factor != 1:
        fir = lowpass_fir(band_center=1 / factor, kernel_length=256, transition_bandwidth=0.03)
    return new_sr, new_hop, factor, fir


# --------------------------------------------------------------------------- #
# Mel filterbank (librosa 0.7 clone in the reference)
# --------------------------------------------------------------------------- #
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = (_MIN_LOG_HZ - 0.0) / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_mel(hz, htk=False):
    """librosa_functions.py:250-298 (scalar or array)."""
    hz = np.asanyarray(hz)
    if htk:
        return 2595.0 * np.log10(1.0 + hz / 700.0)
    mel = (hz - 0.0) / _F_SP
    if hz.ndim:
        hi = hz >= _MIN_LOG_HZ
        mel[hi] = _MIN_LOG_MEL + np.log(hz[hi] / _MIN_LOG_HZ) / _LOGSTEP
    elif hz >= _MIN_LOG_HZ:
        mel = _MIN_LOG_MEL + np.log(hz / _MIN_LOG_HZ) / _LOGSTEP
    return mel


def _mel_to_hz(mel, htk=False):
    """librosa_functions.py:201-247."""
    mel = np.asanyarray(mel)
    if htk:
        return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    hz = 0.0 + _F_SP * mel
    if mel.ndim:
        hi = mel >= _MIN_LOG_MEL
        hz[hi] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mel[hi] - _MIN_LOG_MEL))
    elif mel >= _MIN_LOG_MEL:
        hz = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mel - _MIN_LOG_MEL))
    return hz


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm=1):
    """Slaney (or HTK) triangular filters, float32 ``(n_mels, n_fft//2+1)``
    (librosa_functions.py:375-486).  ``norm == 1`` -> area normalised."""
    if fmax is None:
        fmax = float(sr) / 2
    if norm is not None and norm != 1 and norm != np.inf:
        raise ParameterError("Unsupported norm: {}".format(repr(norm)))
    n_mels = int(n_mels)
    n_freq = int(1 + n_fft // 2)

    fft_hz = np.linspace(0, float(sr) / 2, n_freq, endpoint=True)
    edges_mel = np.linspace(_hz_to_mel(fmin, htk=htk), _hz_to_mel(fmax, htk=htk), n_mels + 2)
    edges_hz = _mel_to_hz(edges_mel, htk=htk)

    widths = np.diff(edges_hz)
    ramps = np.subtract.outer(edges_hz, fft_hz)

    weights = np.zeros((n_mels, n_freq), dtype=np.float32)
    for i in range(n_mels):
        rising = -ramps[i] / widths[i]
        falling = ramps[i + 2] / widths[i + 1]
        weights[i] = np.maximum(0, np.minimum(rising, falling))
    if norm == 1:
        enorm = 2.0 / (edges_hz[2 : n_mels + 2] - edges_hz[:n_mels])
        weights *= enorm[:, np.newaxis]

    if not np.all((edges_hz[:-2] == 0) | (weights.max(axis=1) > 0)):
        warnings.warn(
            "Empty filters detected in mel frequency basis. "
            "Some channels will produce empty responses. "
            "Try increasing your sampling rate (and fmax) or "
            "reducing n_mels."
        )
    return weights


# --------------------------------------------------------------------------- #
# Gammatone filterbank (Ellis / Slaney ERB filters sampled on the FFT grid)
# --------------------------------------------------------------------------- #
def gammatone_filterbank(sr, n_fft, n_bins=64, fmin=20.0, fmax=None):
    """float32 ``(n_bins, n_fft//2+1)`` magnitude responses of 4th-order
    gammatone filters at ERB-spaced centre frequencies, times ``1/n_fft``
    (librosa_functions.py:13-198).  ``width = 1.0``; centre frequencies run
    low -> high."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = int(n_bins)
    maxlen = int(n_fft // 2 + 1)

    ear_q, min_bw, order, gt_order, width = 9.26449, 24.7, 1, 4, 1.0
    em = ear_q * min_bw
    idx = np.array(range(n_bins)) + 1
    cf = (fmax + em) * np.exp(idx * (-np.log(fmax + em) + np.log(fmin + em)) / n_bins) - em
    cf = cf[::-1]

    unit = np.exp(1j * 2 * np.pi * np.array(range(int(n_fft / 2 + 1))) / n_fft)

    erb = width * np.power(np.power(cf / ear_q, order) + np.power(min_bw, order), 1 / order)
    B = 1.019 * 2 * np.pi * erb
    r = np.exp(-B / sr)
    theta = 2 * np.pi * cf / sr
    pole = r * np.exp(1j * theta)
    T = 1 / sr
    ebt = np.exp(B * T)
    cpt = 2 * cf * np.pi * T
    ccpt = 2 * T * np.cos(cpt)
    scpt = 2 * T * np.sin(cpt)
    rp = np.sqrt(3 + 2 ** 1.5)
    rm = np.sqrt(3 - 2 ** 1.5)
    a11 = -np.divide(np.divide(ccpt, ebt) + np.divide(rp * scpt, ebt), 2)
    a12 = -np.divide(np.divide(ccpt, ebt) - np.divide(rp * scpt, ebt), 2)
    a13 = -np.divide(np.divide(ccpt, ebt) + np.divide(rm * scpt, ebt), 2)
    a14 = -np.divide(np.divide(ccpt, ebt) - np.divide(rm * scpt, ebt), 2)
    zeros = -np.array([a11, a12, a13, a14]) / T

    # Gain of the cascade at the centre frequency (Slaney's MakeERBFilters).
    e4 = np.exp(4 * 1j * cf * np.pi * T)
    e2 = np.exp(-(B * T) + 2 * 1j * cf * np.pi * T)
    c2 = np.cos(2 * cf * np.pi * T)
    s2 = np.sin(2 * cf * np.pi * T)
    q_m = np.sqrt(3 - 2 ** (3 / 2))
    q_p = np.sqrt(3 + 2 ** (3 / 2))

    def term(sign, q):
        return -2 * e4 * T + 2 * e2 * T * (c2 + sign * q * s2)

    gain = np.abs(
        term(-1, q_m)
        * term(+1, q_m)
        * term(-1, q_p)
        * term(+1, q_p)
        / (-2 / np.exp(2 * B * T) - 2 * e4 + 2 * (1 + e4) / np.exp(B * T)) ** 4
    )

    col = lambda v: np.reshape(v, (n_bins, 1))
    wts = np.zeros([n_bins, n_fft], dtype=np.float32)
    wts[:, : int(n_fft / 2 + 1)] = (
        ((T ** 4) / col(gain))
        * np.abs(unit - col(zeros[0]))
        * np.abs(unit - col(zeros[1]))
        * np.abs(unit - col(zeros[2]))
        * np.abs(unit - col(zeros[3]))
        * (np.abs(np.power(np.multiply(col(pole) - unit, np.conj(col(pole)) - unit), -gt_order)))
    )
    return (1 / n_fft) * wts[:, :maxlen]


# --------------------------------------------------------------------------- #
# DCT-II (ortho) as a dense matrix
# --------------------------------------------------------------------------- #
def dct2_ortho_matrix(n_out: int, n_in: int) -> np.ndarray:
    """Rows 0..n_out-1 of the orthonormal DCT-II of size ``n_in`` as float32.

    The reference computes the same transform through an FFT of the even/odd
    re-ordered input plus a twiddle (mel.py:281-307): V_k = 2*Re(FFT(v)_k *
    e^{-i*pi*k/(2N)}), scaled by 1/(2*sqrt(N)) for k=0 and 1/(2*sqrt(N/2))
    otherwise, i.e. D[k, n] = c_k * cos(pi*(2n+1)*k/(2N)) with c_0 = 1/sqrt(N),
    c_k = sqrt(2/N).
    """
    n = np.arange(n_in, dtype=np.float64)
    k = np.arange(n_out, dtype=np.float64)[:, None]
    mat = np.cos(np.pi * (2 * n[None, :] + 1) * k / (2 * n_in))
    mat[0] *= 1.0 / math.sqrt(n_in)
    mat[1:] *= math.sqrt(2.0 / n_in)
    return mat.astype(np.float32)


# --------------------------------------------------------------------------- #
# Combined frequency / periodicity representation (cfp.py)
# --------------------------------------------------------------------------- #
def blackmanharris_window(window_size: int) -> np.ndarray:
    """cfp.py:89 — ``scipy.signal.blackmanharris(window_size)`` (symmetric, 4-term), float64.
    (SciPy >= 1.13 only keeps the function under ``scipy.signal.windows``.)"""
    return _sps.windows.blackmanharris(int(window_size))


def cfp_axes(fr, fs, fc, tc):
    """cfp.py:83-106 — the transform size and the crops of a CFP module:
    ``N = int(fs / fr)``, the frequency axis ``f`` (first ``HighFreqIdx`` points of
    ``fs * linspace(0, 0.5, N // 2)``), the quefrency axis ``q = arange(HighQuefIdx) / fs``, and
    the two cut-off indices of the non-linearities."""
    N = int(fs / float(fr))
    f_full = fs * np.linspace(0, 0.5, np.round(N // 2), endpoint=True)
    tc_idx = round(fs * tc)
    fc_idx = round(fc / fr)
    high_freq = int(round((1 / tc) / fr) + 1)
    high_quef = int(round(fs / fc) + 1)
    return dict(N=N, f=f_full[:high_freq], q=np.arange(high_quef) / float(fs), tc_idx=tc_idx,
                fc_idx=fc_idx, HighFreqIdx=high_freq, HighQuefIdx=high_quef)


def _triangle_rows(axis_hz: np.ndarray, lo: int, hi: int, left: float, centre: float, right: float,
                   row: np.ndarray) -> None:
    """One band of a log-frequency matrix: points of ``axis_hz[lo:hi]`` strictly inside
    (left, centre) get the rising edge, strictly inside (centre, right) the falling edge
    (cfp.py:219-227 / 233-244; points equal to a centre stay 0, as in the reference)."""
    if hi > axis_hz.shape[0]:  # the reference indexes point by point and fails on the first one past the end
        raise IndexError(f"index {axis_hz.shape[0]} is out of bounds for axis 0 with size {axis_hz.shape[0]}")
    seg = axis_hz[lo:hi]
    rising = (seg > left) & (seg < centre)
    falling = (seg > centre) & (seg < right)
    vals = np.zeros(seg.shape[0], dtype=np.float64)
    with np.errstate(invalid="ignore"):
        vals[rising] = (seg[rising] - left) / (centre - left)
        vals[falling] = (right - seg[falling]) / (right - centre)
    row[lo:hi] = vals


def cfp_logfreq_matrices(f: np.ndarray, q: np.ndarray, fr, fc, tc, NumPerOct, fs):
    """cfp.py:195-246 ``create_logfreq_matrix`` — triangular maps from the linear frequency axis
    ``f`` and from the quefrency axis ``q`` (as pitch ``1 / q``) onto ``NumPerOct`` bands per
    octave between ``fc`` and ``1 / tc``; float64 ``(Nest - 1, len(f))`` and ``(Nest - 1, len(q))``.
    Rows 0 and Nest - 2 stay empty, as in the reference (its loops run over the interior bands)."""
    start, stop = fc, 1 / tc
    n_est = int(np.ceil(np.log2(stop / start)) * NumPerOct)
    centres = []
    for i in range(0, n_est):
        c = start * pow(2, float(i) / NumPerOct)
        if c < stop:
            centres.append(c)
        else:
            break
    n_est = len(centres)
    freq_map = np.zeros((n_est - 1, len(f)), dtype=np.double)
    for i in range(1, n_est - 1):
        lo = int(round(centres[i - 1] / fr))
        hi = int(round(centres[i + 1] / fr) + 1)
        if lo >= hi - 1:
            freq_map[i, lo] = 1  # the band is narrower than one bin
        else:
            _triangle_rows(f, lo, hi, centres[i - 1], centres[i], centres[i + 1], freq_map[i])
    with np.errstate(divide="ignore"):
        pitch = 1 / q  # q[0] = 0 -> inf, never inside a band
    quef_map = np.zeros((n_est - 1, len(pitch)), dtype=np.double)
    for i in range(1, n_est - 1):
        lo = int(round(fs / centres[i + 1]))
        hi = int(round(fs / centres[i - 1]) + 1)
        _triangle_rows(pitch, lo, hi, centres[i - 1], centres[i], centres[i + 1], quef_map[i])
    return freq_map, quef_map
