set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
python tools/batch_stages.py --tag default > gpurun_out/r2c_stages.json 2> gpurun_out/r2c_stages.err
python tools/batch_stages.py --tag tum --shape 480 640 >> gpurun_out/r2c_stages.json 2>> gpurun_out/r2c_stages.err
cat gpurun_out/r2c_stages.json; tail -3 gpurun_out/r2c_stages.err
python tools/single_frame_stages.py > gpurun_out/r2c_single.log 2>&1; cat gpurun_out/r2c_single.log
( time python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2>&1 | tail -4
tail -5 gpurun_out/r2c_bench.err; cut -c1-1500 gpurun_out/r2c_bench.json
( time python bench.py --impl reference > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err ) 2>&1 | tail -4
cut -c1-1500 gpurun_out/r2c_bench_ref.json
ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_r2_c python tools/prof_extract.py --batch 256 --iters 1 > gpurun_out/ncu_r2_c.log 2>&1
tail -2 gpurun_out/ncu_r2_c.log
