"""Several GPUs driven from ONE process (-m gpu, needs >= 2 devices): the reference's only multi-GPU
mode is ``torch.nn.DataParallel`` (Installation/tests/test_stft.py:122-141, test_cqt.py:271-292) --
module replicas called concurrently from one Python thread per GPU.  The kernels' per-device launch
attributes, the per-device packed-basis caches and the layout registries must all cope."""
import pytest
import torch

from helpers import build

pytestmark = pytest.mark.gpu

needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 CUDA devices")

CASES = [
    ("STFT", dict(n_fft=2048, hop_length=512, output_format="Magnitude"), (8, 40000)),
    ("MelSpectrogram", dict(sr=22050, n_fft=2048, hop_length=512, n_mels=128), (8, 40000)),
    ("MFCC", dict(sr=16000), (8, 32000)),
    ("CQT1992v2", dict(sr=22050, n_bins=60, fmin=110.0), (6, 44100)),
    ("CQT2010v2", dict(sr=22050, n_bins=84), (6, 44100)),
]


@needs2
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_dataparallel_matches_single_device(case):
    cls, ctor, shape = case
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(3))
    mod = build(cls, ctor).cuda(0)
    with torch.no_grad():
        want = mod(x.cuda(0))
        dp = torch.nn.DataParallel(mod, device_ids=[0, 1])
        got = dp(x.cuda(0))
        got2 = dp(x.cuda(0))   # second call: caches populated on both devices
    assert got.shape == want.shape
    assert torch.equal(got, want) and torch.equal(got2, want)


@needs2
def test_same_module_type_on_two_devices_sequentially():
    """A module on cuda:1 after one on cuda:0 (first > 48 KB dynamic-smem launch per device)."""
    x = torch.randn(4, 30000, generator=torch.Generator().manual_seed(5))
    outs = []
    for d in (0, 1):
        m = build("MelSpectrogram", dict(sr=22050, n_fft=1024, hop_length=256, n_mels=64)).cuda(d)
        with torch.no_grad():
            outs.append(m(x.cuda(d)).cpu())
    assert torch.equal(outs[0], outs[1])
