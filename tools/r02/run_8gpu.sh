set -x
nvidia-smi -L > gpurun_out/r02_8gpu_smi.txt
for n in 8 4 2; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n"
$TR bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/r02_bench_${n}gpu.err; echo "rc=$?"
done
python bench.py --steps 20 --warmup 5 --no-c3 > gpurun_out/r02_bench_fanout_8gpu.json 2> gpurun_out/r02_bench_fanout_8gpu.err; echo "rc=$?"
python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_gputests_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_8gpu.log
