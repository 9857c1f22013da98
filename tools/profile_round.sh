set -x
python bench.py > gpurun_out/bench_r1_k.json 2> gpurun_out/bench_r1_k.err
python bench.py --impl reference > gpurun_out/bench_r1_k_ref.json 2> gpurun_out/bench_r1_k_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --knn-steps 1 > gpurun_out/ncu_l_k.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_r1_k python tools/prof_extract.py --batch 256 --iters 1 > gpurun_out/ncu_k.log 2>&1
ls -la gpurun_out | tail -8
