python -m pytest tests/test_gpu_matcher_host.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2s_pytest.log
