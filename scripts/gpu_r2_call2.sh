mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest2.log
tail -5 gpurun_out/pytest2.log
timeout 900 python scripts/ab_exact.py > gpurun_out/ab_exact.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_exact.log
