set -x
python tools/r02/time_a.py > gpurun_out/r02_time_a.txt 2>&1
