set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -25 gpurun_out/r2e_pytest.log
python tools/batch_stages.py --tag default_lanes2 > gpurun_out/r2e_stages.json 2> gpurun_out/r2e_stages.err
ORBB_LANES=3 python tools/batch_stages.py --tag lanes3 >> gpurun_out/r2e_stages.json 2>> gpurun_out/r2e_stages.err
ORBB_LANES=4 python tools/batch_stages.py --tag lanes4 >> gpurun_out/r2e_stages.json 2>> gpurun_out/r2e_stages.err
ORBB_LANES=4 ORBB_LANES_MIN=32 python tools/batch_stages.py --tag lanes4_tum --shape 480 640 >> gpurun_out/r2e_stages.json 2>> gpurun_out/r2e_stages.err
python tools/batch_stages.py --tag tum --shape 480 640 >> gpurun_out/r2e_stages.json 2>> gpurun_out/r2e_stages.err
cat gpurun_out/r2e_stages.json; tail -3 gpurun_out/r2e_stages.err
python bench.py --no-knn --no-stereo --no-shapes --no-config3 --no-cpu-baseline > gpurun_out/r2e_bench_l2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_l2.json')); print('lanes host off: value',d['value'],'e2e',d['e2e']['value'],d['e2e']['copy_ceiling_frames_per_s'])"
ORBB_LANES_HOST=1 python bench.py --no-knn --no-stereo --no-shapes --no-config3 --no-cpu-baseline > gpurun_out/r2e_bench_lh.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_lh.json')); print('lanes host on: value',d['value'],'e2e',d['e2e']['value'])"
