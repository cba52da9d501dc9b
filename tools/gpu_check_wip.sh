#!/usr/bin/env bash
# First GPU pass over the experimental kernels of this branch (each step bounded; the kernels' mbarrier
# waits trap after 4 s, so a wrong pipeline shows up as a CUDA error, not a hang).
set -u
mkdir -p gpurun_out
run() { echo "== $*"; "$@" 2>&1 | tail -6; }
# 1. radix-2 STFT family
NNAB_RADIX=2 run timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 \
    -k "stft_cfg1 or stft_2048 or stft_constant_pad or mel_ or mfcc_ or gammatone"
for wl in stft2048 cfg2 cfg5; do
  NNAB_RADIX=2 timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/wip_radix2_$wl.json 2>> gpurun_out/wip_err.txt
  cut -c1-260 gpurun_out/wip_radix2_$wl.json
done
# 1a. tile-width A/B: 128 bins per tile, N = 256 per MMA, TMEM single-buffered
for wl in stft2048 cfg2; do
  NNAB_RADIX=2 NNAB_RADIX_BN=256 timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/wip_radix2_bn256_$wl.json 2>> gpurun_out/wip_err.txt
  cut -c1-260 gpurun_out/wip_radix2_bn256_$wl.json
done
NNAB_RADIX=2 NNAB_RADIX_BN=256 run timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "stft_2048 or mel_"
# 1c. radix 4 (64-column segments, add/swap butterflies)
NNAB_RADIX=4 run timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "stft_cfg1 or stft_2048 or mel_ or mfcc_"
for wl in stft2048 cfg2 cfg5; do
  NNAB_RADIX=4 timeout 150 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/wip_radix4_$wl.json 2>> gpurun_out/wip_err.txt
  cut -c1-260 gpurun_out/wip_radix4_$wl.json
done
# 1b. the same through the host layer's own selection (structure check + explicit layout request)
NNAUDIO_B200_EXPERIMENTAL=1 run timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 120
# 2. per-K-block width, CQT1992v2
NNAB_VARN=1 run timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 \
    -k "cqt1992v2 or sweep-cqt-1992"
NNAB_VARN=1 timeout 150 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/wip_varn_cfg3.json 2>> gpurun_out/wip_err.txt
cut -c1-260 gpurun_out/wip_varn_cfg3.json
# 3. dedicated FIR stage in the pyramid's training path
NNAUDIO_B200_DECIM_BWD=fir run timeout 150 python -m pytest tests/test_backward.py -m gpu -q --timeout 100 -k "cqt2010 or vqt"
NNAUDIO_B200_DECIM_BWD=fir run timeout 100 python tools/bench_training.py --iters 10 --only cqt
# 4. 4-term split: forward error of the training path
NNAB_SPLIT4=1 run timeout 150 python -m pytest tests/test_backward.py -m gpu -q --timeout 100 -k "stft or mel"
tail -5 gpurun_out/wip_err.txt
