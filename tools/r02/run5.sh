set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_c.json 2> gpurun_out/r02_bench_1gpu_c.err; echo "bench rc=$?"
