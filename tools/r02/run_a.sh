set -x
GPBT_B200_LIB=$PWD/build/head/libgpbt_head.so python tools/r02/time_a.py > gpurun_out/r02_time_a_head.txt 2>&1
python tools/r02/time_a.py > gpurun_out/r02_time_a.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_a_tests.log
