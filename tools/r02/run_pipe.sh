set -x
python tools/r02/pipe_timing.py 4096 > gpurun_out/r02_pipe_timing.txt 2>&1; echo rc=$?
python tools/r02/pipe_timing.py 8192 >> gpurun_out/r02_pipe_timing.txt 2>&1; echo rc=$?
