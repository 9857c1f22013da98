set -x
python -m pytest tests/test_gpu_extract.py -m gpu -x -q -k "not switch" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
python tools/batch_stages.py --tag desc16 > gpurun_out/r2k_stages.json 2> gpurun_out/r2k_stages.err
ORBB_BLUR_UNROLL=1 python tools/batch_stages.py --tag blur_unroll4 >> gpurun_out/r2k_stages.json 2>> gpurun_out/r2k_stages.err
ORBB_BLUR_UNROLL=1 python tools/batch_stages.py --tag blur_unroll4_b >> gpurun_out/r2k_stages.json 2>> gpurun_out/r2k_stages.err
python tools/batch_stages.py --tag desc16_b >> gpurun_out/r2k_stages.json 2>> gpurun_out/r2k_stages.err
cat gpurun_out/r2k_stages.json; tail -3 gpurun_out/r2k_stages.err
