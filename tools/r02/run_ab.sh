set -x
nvidia-smi --query-gpu=name,serial,uuid --format=csv > gpurun_out/r02_ab.txt
for rep in 1 2 3; do
python tools/r02/ab_timing.py 4096 "chol_pipe=1" "chol_pipe=1,chol_lag=0" "" "chol_pipe=2" >> gpurun_out/r02_ab.txt 2>&1
done
python tools/r02/ab_timing.py 8192 "chol_pipe=1" "chol_pipe=1,chol_lag=0" "" "chol_pipe=3" >> gpurun_out/r02_ab.txt 2>&1
python tools/r02/ab_timing.py 2048 "chol_pipe=1" "chol_pipe=1,chol_lag=0" "" >> gpurun_out/r02_ab.txt 2>&1
