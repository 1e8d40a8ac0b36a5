# final one-GPU pass of the round: smoke, GPU tests, the bench (both arms), launch lists, full ncu captures
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.txt
python -m pytest tests -m gpu -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_g.json 2> gpurun_out/r02_bench_1gpu_g.err ) 2> gpurun_out/r02_bench_time.txt
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2>> gpurun_out/r02_bench_time.txt
# launch list of a short bench run (the same command as the bench, fewer steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-configs --no-dgemm --dense-steps 1 --sustained-s 0 --cpu-rows 0 > gpurun_out/r02_ncu_bench.log 2>&1
# launch list of one dense-path call (N = 4096, default streams)
REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chol_fused|pc_predict|backtransform|bounds' --csv --log-file gpurun_out/r02_dense_launches.csv python tools/r02/profile_dense.py 4096 > gpurun_out/r02_ncu_dl.log 2>&1
# full captures (single stream, 2048 walkers): first, two mid and the last two panel launches; one factor launch
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_fused_panel --launch-skip 11 --launch-count 9 \
  -o gpurun_out/r02_chol_fused_panel -f python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_f1.log 2>&1
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_fused_factor --launch-skip 15 --launch-count 1 \
  -o gpurun_out/r02_chol_fused_factor -f python tools/r02/profile_dense.py 2048 fused 100000 > gpurun_out/r02_ncu_f2.log 2>&1
