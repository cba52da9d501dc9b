#!/usr/bin/env bash
# Last GPU call of the round: newest tests first, then the rest of the GPU suite, smoke, default bench.
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_griffin_lim.py tests/test_backward.py tests/test_istft.py -m gpu -q \
    --timeout 120 > gpurun_out/fin_new.log 2>&1
tail -4 gpurun_out/fin_new.log
timeout 300 python -m pytest tests -m gpu -q --timeout 120 --deselect tests/test_griffin_lim.py \
    --deselect tests/test_backward.py --deselect tests/test_istft.py > gpurun_out/fin_rest.log 2>&1
tail -4 gpurun_out/fin_rest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.txt 2>&1
tail -3 gpurun_out/fin_smoke.txt
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/fin_bench_cfg2.json 2> gpurun_out/fin_err.txt
cut -c1-400 gpurun_out/fin_bench_cfg2.json
