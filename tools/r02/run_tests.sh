set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.txt
python tools/r02/fused_verify.py 300 > gpurun_out/r02_fused_verify.txt 2>&1; echo rc=$?
