set -x
python tools/r02/fused_verify.py 1200 > gpurun_out/r02_fused_verify.txt 2>&1; echo rc=$?
