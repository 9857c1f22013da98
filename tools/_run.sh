set -x
python -m pytest tests/test_gpu_extract.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
python tools/latency_probe.py > gpurun_out/r2h_lat.log 2>&1; cat gpurun_out/r2h_lat.log
ORBB_BRANCH_FRAMES=0 python tools/latency_probe.py > gpurun_out/r2h_lat_nobranch.log 2>&1; cat gpurun_out/r2h_lat_nobranch.log
ORBB_GRAPH_NO_PDL=1 python tools/latency_probe.py > gpurun_out/r2h_lat_nopdl.log 2>&1; cat gpurun_out/r2h_lat_nopdl.log
python tools/batch_stages.py --tag default > gpurun_out/r2h_stages.json 2>&1; cat gpurun_out/r2h_stages.json
