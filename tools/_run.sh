set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -5 gpurun_out/r2p_pytest.log
python tools/batch_stages.py --tag final > gpurun_out/r2p_stages.json 2> gpurun_out/r2p_stages.err
python tools/batch_stages.py --tag final_tum --shape 480 640 >> gpurun_out/r2p_stages.json 2>> gpurun_out/r2p_stages.err
cat gpurun_out/r2p_stages.json | cut -c1-330; tail -3 gpurun_out/r2p_stages.err
python tools/latency_probe.py > gpurun_out/r2p_latency.log 2>&1; tail -7 gpurun_out/r2p_latency.log
