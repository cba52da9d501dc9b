"""Parity at BASELINE.json's FULL sizes (-m gpu).

The oracle cannot run whole batches in seconds, so every configuration runs on the GPU at
its real (B, L) and a few clips (first / middle / last) are compared with the fp64 oracle
computed per clip — this exercises the batch-wide virtual-frame indexing, the tile tails
and the per-clip MFCC floor at scale — plus size-independent properties (linearity of the
complex transforms, clip permutation equivariance)."""
import warnings

import pytest
import torch

from helpers import build, rel_errors, run_oracle

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (class, ctor, B, L, forward kwargs, tolerance)
    "cfg1": ("STFT", dict(n_fft=512, hop_length=256, sr=16000), 1, 16000, dict(output_format="Complex"), 1e-4),
    "cfg2": ("MelSpectrogram", dict(sr=22050, n_fft=2048, hop_length=512, n_mels=128), 64, 220500, {}, 1e-4),
    "cfg3": ("CQT1992v2", dict(sr=44100, n_bins=84, bins_per_octave=12, fmin=32.7), 128, 441000,
             dict(output_format="Magnitude"), 1e-4),
    "cfg4": ("CQT2010v2", dict(sr=22050, n_bins=88), 256, 661500, dict(output_format="Magnitude"), 1e-4),
    "cfg5": ("MFCC", dict(sr=16000), 1024, 80000, {}, 1e-4),
}


def _noise(B, L, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(B, L, generator=g, device="cuda", dtype=torch.float32)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_batch_sampled_clips_match_oracle(name):
    cls, ctor, B, L, kw, tol = CONFIGS[name]
    mod = build(cls, ctor).cuda()
    x = _noise(B, L, 1234)
    if cls == "MFCC":  # different levels per clip: the top_db floor must be per clip
        x = x * torch.logspace(0, -3, B, device="cuda")[:, None]
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = mod(x, **kw)
    torch.cuda.synchronize()
    hop = ctor.get("hop_length", 512)
    assert y.shape[0] == B and y.shape[2] == L // hop + 1
    assert torch.isfinite(y).all()
    for b in sorted({0, B // 2, B - 1}):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = run_oracle(cls, mod, x[b:b + 1].cpu().numpy(), kw)
        emax, el2 = rel_errors(y[b:b + 1].cpu().numpy(), ref)
        assert emax < tol and el2 < tol, (name, b, emax, el2)


def test_stft_linearity_and_clip_permutation_at_full_size():
    """STFT 'Complex' is linear in x and acts on clips independently."""
    mod = build("STFT", dict(n_fft=2048, hop_length=512, output_format="Complex")).cuda()
    B, L = 64, 220500
    x, z = _noise(B, L, 1), _noise(B, L, 2)
    with torch.no_grad():
        a = mod(x)
        b = mod(z)
        c = mod(0.5 * x - 2.0 * z)
        perm = torch.randperm(B, device="cuda")
        d = mod(x[perm])
    lin = 0.5 * a - 2.0 * b
    scale = lin.abs().max()
    assert ((c - lin).abs().max() / scale).item() < 1e-4
    assert torch.equal(d, a[perm]), "a clip's spectrogram must not depend on its position in the batch"


def test_mel_fused_and_unfused_paths_agree_at_full_size():
    """The fused tensor-core epilogue vs the CUDA-core filterbank GEMM on the full cfg2 batch."""
    import os
    mod = build("MelSpectrogram", dict(sr=22050, n_fft=2048, hop_length=512, n_mels=128)).cuda()
    x = _noise(64, 220500, 3)
    with torch.no_grad():
        fused = mod(x)
        os.environ["NNAUDIO_B200_PATH"] = "simt"
        try:
            plain = mod(x)
        finally:
            os.environ.pop("NNAUDIO_B200_PATH", None)
    emax, el2 = rel_errors(fused.cpu().numpy(), plain.cpu().numpy())
    assert emax < 1e-4 and el2 < 1e-4, (emax, el2)
