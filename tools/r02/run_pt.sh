set -x
python -m pytest tests/test_gpu_ptlmc.py -m gpu -x -q > gpurun_out/r02_pt_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pt_tests.log
python bench.py --steps 5 --warmup 3 --no-dgemm --dense-steps 0 --sustained-s 0 --cpu-rows 0 --no-fanout > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; echo rc=$?
