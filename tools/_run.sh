python -m pytest tests/test_gpu_extract.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
echo kpw2; python tools/latency_probe.py 2>&1 | tail -3
echo kpw1; ORBB_LIB=orb_slam3_ros_b200/liborbb200_kpw1.so python tools/latency_probe.py 2>&1 | tail -3
echo kpw4; ORBB_LIB=orb_slam3_ros_b200/liborbb200_kpw4.so python tools/latency_probe.py 2>&1 | tail -3
