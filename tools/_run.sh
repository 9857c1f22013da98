python -m pytest tests/test_gpu_matcher_host.py -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2r_pytest.log
python bench.py --no-cpu-baseline --no-config3 --no-shapes --no-stereo --no-knn --no-matcher-rows --sustain-s 0 --steps 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -3 gpurun_out/r2r_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print({k:v for k,v in d.items() if 'single' in k}, d['value'])
P
