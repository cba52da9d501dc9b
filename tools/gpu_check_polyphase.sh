#!/usr/bin/env bash
# Pyramid training path after the polyphase decimation adjoint: parity + timing, both kernel families.
set -u
mkdir -p gpurun_out
for fam in simt tc; do
  NNAUDIO_B200_DECIM_BWD=$fam timeout 120 python -m pytest tests/test_backward.py -m gpu -q --timeout 100 \
      -k "cqt2010 or vqt" > gpurun_out/poly_${fam}.log 2>&1
  tail -2 gpurun_out/poly_${fam}.log
  cp gpurun_out/backward_errors_auto.json gpurun_out/poly_errors_${fam}.json 2>/dev/null
  NNAUDIO_B200_DECIM_BWD=$fam timeout 100 python tools/bench_training.py --iters 10 --only cqt 2>&1 | tail -3
  cp gpurun_out/bench_training.json gpurun_out/bench_training_${fam}.json 2>/dev/null
done
