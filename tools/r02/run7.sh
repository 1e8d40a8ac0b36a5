set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.txt
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_e.json 2> gpurun_out/r02_bench_1gpu_e.err ) 2> gpurun_out/r02_bench_time.txt
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>> gpurun_out/r02_bench_time.txt
