#!/bin/bash
# One round's measurement set on a GPU box (run under gpurun, one GPU):  bash tools/profile_round.sh <tag>
#   bench line, reference arm, the ncu launch list of the bench command (short legs), one `ncu --set full` capture of a 256-frame batch
#   (extraction kernels, with source correlation) and one of the matcher-side kernels (stereo, scans, BoW, 2-NN).  Every ncu pass runs
#   under its own `timeout`; the reports stay below gpurun's 64 MiB return limit (kernel filter on the matcher capture).
# Outputs land in gpurun_out/; summarise them here with
#   python tools/summarize_profiles.py <tag> --rep gpurun_out/prof_<tag>.ncu-rep --launches gpurun_out/launches_<tag>.csv
#   python tools/summarize_profiles.py <tag>_matcher --rep gpurun_out/prof_<tag>_matcher.ncu-rep
#   python tools/ncu_lines.py gpurun_out/prof_<tag>.ncu-rep k_fast_cell --csv profiles/<tag>_fast_cell_lines.csv
TAG=${1:-r2}
set -x
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --knn-steps 1 --sustain-s 0 --no-config3 --no-matcher-rows --latency-reps 2 --no-shapes --stereo-steps 1 > gpurun_out/ncu_l_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_${TAG} python tools/prof_extract.py --batch 256 --iters 1 > gpurun_out/ncu_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'k_best2_csr|k_bow_|k_distinctive|k_knn2_|k_ratio_test|k_remap|k_rotation_check|k_search_area_topk|k_stereo_|k_undistort' -c 40 -f -o gpurun_out/prof_${TAG}_matcher python tools/prof_matcher.py --knn --knn-nq 50000 > gpurun_out/ncu_${TAG}_matcher.log 2>&1
ls -la gpurun_out | tail -8
