set -x
# launch list of a short bench run (the same command as the bench, fewer steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-configs --no-dgemm --dense-steps 1 --sustained-s 0 --cpu-rows 0 > gpurun_out/r02_ncu_bench.log 2>&1
# launch list of one dense-path call (N = 4096, default streams)
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chol_fused|pc_predict|backtransform|bounds' --csv --log-file gpurun_out/r02_dense_launches.csv python tools/r02/profile_dense.py 4096 > gpurun_out/r02_ncu_dl.log 2>&1
# full captures: two mid panels, the factor kernel, kernel (a)
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_fused_panel --launch-skip 14 --launch-count 2 \
  -o gpurun_out/r02_chol_fused_panel -f python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_f1.log 2>&1
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_fused_factor --launch-skip 15 --launch-count 1 \
  -o gpurun_out/r02_chol_fused_factor -f python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_f2.log 2>&1
