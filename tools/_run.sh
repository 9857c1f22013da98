set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
python tools/single_frame_stages.py > gpurun_out/r2i_single.log 2>&1; cat gpurun_out/r2i_single.log
python tools/latency_probe.py > gpurun_out/r2i_latency.log 2>&1; tail -7 gpurun_out/r2i_latency.log
ORBB_GRAPH_PDL=1 python tools/latency_probe.py > gpurun_out/r2i_latency_pdl.log 2>&1; tail -7 gpurun_out/r2i_latency_pdl.log
ORBB_FAST_LATENCY_FRAMES=0 python tools/latency_probe.py > gpurun_out/r2i_latency_1w.log 2>&1; tail -7 gpurun_out/r2i_latency_1w.log
( time timeout 600 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ) 2>&1 | tail -4
tail -3 gpurun_out/r2i_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','single_frame_latency_ms','single_frame_c_abi_ms','single_frame_with_pyramid_ms','sustained','gpu_launches')})
print(d['e2e']); print(d['cpu_baseline']); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['stage_ms'])
P
