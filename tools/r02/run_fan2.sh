set -x
python -m pytest tests/test_multi_gpu.py -m gpu -q -k fanout > gpurun_out/r02_gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_2gpu.log
python bench.py --steps 10 --warmup 3 --no-c3 --no-dgemm --dense-steps 0 --sustained-s 0 --cpu-rows 0 > gpurun_out/r02_bench_fanout_2gpu_b.json 2> gpurun_out/r02_bench_fanout_2gpu_b.err; echo "rc=$?"
