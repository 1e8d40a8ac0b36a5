set -x
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chol_fused|pc_predict|backtransform|bounds' --csv --log-file gpurun_out/r02_fused_launches.csv python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_fl.log 2>&1
python tools/r02/cf_timing.py 2048 > gpurun_out/r02_cf_timing.txt 2>&1
