set -x
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1; head -12 gpurun_out/r2f_topo.txt
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2f_pytest_multi.log 2>&1; tail -4 gpurun_out/r2f_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; tail -3 gpurun_out/r2f_bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2f_bench_n2.json'))
print('N=2 value',d['value'],'e2e',d['e2e'],'\nknn',{k:d['knn'][k] for k in ('value','result_crc32','matched_ratio_0.7','sharding')},'\nconfig3',d['config3'],'\nhost',d['host'])
P
python bench.py --no-stereo --no-shapes --no-cpu-baseline --no-matcher-rows > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print('N=1 value',d['value'],'e2e',d['e2e']['value'],d['e2e']['copy_ceiling_frames_per_s'],'knn crc',d['knn']['result_crc32'],'config3',d['config3']['value'],d['config3']['e2e'])
P
