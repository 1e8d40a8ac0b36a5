set -x
nvidia-smi -L > gpurun_out/r02_2gpu_smi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu_fused.json 2> gpurun_out/r02_bench_2gpu_fused.err; echo "rc=$?"
$TR bench.py --gpus 2 --steps 20 --warmup 5 --clock-sampler off > gpurun_out/r02_bench_2gpu_fused_noclk.json 2> gpurun_out/r02_bench_2gpu_fused_noclk.err; echo "rc=$?"
$TR bench.py --gpus 2 --steps 20 --warmup 5 --collective nccl > gpurun_out/r02_bench_2gpu_nccl.json 2> gpurun_out/r02_bench_2gpu_nccl.err; echo "rc=$?"
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_2gpu.log
python bench.py --steps 10 --warmup 3 --no-c3 > gpurun_out/r02_bench_fanout_2gpu.json 2> gpurun_out/r02_bench_fanout_2gpu.err; echo "rc=$?"
