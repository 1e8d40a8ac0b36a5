set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or cholesky or dense" > gpurun_out/r02_ab_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_ab_tests.log
for rep in 1 2 3; do
GPBT_B200_LIB=$PWD/build/head/libgpbt_head.so python tools/r02/ab_timing.py 4096 "" >> gpurun_out/r02_ab.txt 2>&1
python tools/r02/ab_timing.py 4096 "" >> gpurun_out/r02_ab.txt 2>&1
done
REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'chol_fused' --csv --log-file gpurun_out/r02_fused_launches.csv python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_fl.log 2>&1
python tools/r02/fused_verify.py 100 > gpurun_out/r02_fused_verify.txt 2>&1
