mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_code" > gpurun_out/pytest_long.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_long.log
tail -5 gpurun_out/pytest_long.log
bash scripts/n2_check.sh
