python tools/batch_stages.py --tag default 2>&1 | tail -1
python tools/batch_stages.py --tag tum --shape 480 640 2>&1 | tail -1
python -m pytest tests/test_gpu_extract.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
