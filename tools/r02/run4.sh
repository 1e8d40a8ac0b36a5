set -x
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_fused_panel --launch-skip 11 --launch-count 5 \
  -o gpurun_out/r02_chol_fused_v3 -f python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_fused.log 2>&1
