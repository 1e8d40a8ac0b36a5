#!/bin/bash
# round-2 baseline on one GPU: full ncu capture of the stepped Cholesky chain (N = 1024 dense call),
# kernel (a) and kernel (b); dense sweep; plain bench
set -x
mkdir -p gpurun_out
python tools/dense_sweep.py > gpurun_out/r02_dense_sweep_base.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_base.json 2> gpurun_out/r02_bench_base.err
# dense call: self check (N=16: 19 chol launches) then N=1024 (19 launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_step --launch-skip 19 --launch-count 19 \
  -o gpurun_out/r02_chol_stepped_base -f python tools/profile_step.py 1 1024 > gpurun_out/r02_ncu_chol.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pc_predict --launch-skip 2 --launch-count 1 \
  -o gpurun_out/r02_pc_predict_base -f python tools/profile_step.py 3 0 > gpurun_out/r02_ncu_a.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_smi.txt
