#!/bin/bash
# one GPU: GPU tests, bench, ncu of the stepped Cholesky chain
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_a.json 2> gpurun_out/r02_bench_1gpu_a.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_step --launch-skip 0 --launch-count 19 \
  -o gpurun_out/r02_chol_stepped_base -f python tools/profile_step.py 1 1024 > gpurun_out/r02_ncu_chol.log 2>&1
