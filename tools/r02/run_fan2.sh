set -x
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "rc=$?"
