#!/usr/bin/env bash
# Short GPU check of the training paths and the v1 CQT modules (under gpurun).
set -u
mkdir -p gpurun_out
K='cqt1992_ or cqt2010_'
timeout 300 python -m pytest tests/test_backward.py tests/test_istft.py -m gpu -q --timeout 120 \
    > gpurun_out/new_backward.log 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "$K" \
    > gpurun_out/new_v1.log 2>&1
# same training cases with the CUDA-core forward kernels (isolates forward-kernel issues)
NNAUDIO_B200_PATH=simt timeout 200 python -m pytest tests/test_backward.py -m gpu -q --timeout 120 \
    -k "cqt2010 or vqt or cqt1992" > gpurun_out/new_backward_simt.log 2>&1
timeout 120 python bench.py --workload stft2048 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/new_bench_stft2048.json 2> gpurun_out/new_err.txt
tail -15 gpurun_out/new_backward.log; tail -8 gpurun_out/new_v1.log; tail -8 gpurun_out/new_backward_simt.log
cat gpurun_out/new_bench_stft2048.json | cut -c1-300
