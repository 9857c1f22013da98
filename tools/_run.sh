set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2o_pytest_multi.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r2o_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print('N=8 value',d['value'],'e2e',d['e2e']['value'],d['e2e'].get('copy_ceiling_frames_per_s'),'\nknn',{k:d['knn'].get(k) for k in ('value','result_crc32','sharding')},'\nconfig3',d.get('config3'))
P
