set -x
python tools/r02/ptlmc_small.py > gpurun_out/r02_ptlmc_small.txt 2>&1
