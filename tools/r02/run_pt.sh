set -x
python -m pytest tests/test_gpu_ptlmc.py -m gpu -x -q > gpurun_out/r02_pt_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pt_tests.log
