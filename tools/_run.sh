python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log
python tools/latency_probe.py 2>&1 | tail -6
ORBB_OCTREE_NO_SMEM=1 python tools/latency_probe.py 2>&1 | tail -3
python tools/batch_stages.py --tag default 2>&1 | tail -2
