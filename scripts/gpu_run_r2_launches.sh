#!/bin/bash
# The launch list of the default bench step alone (my kernels only):  bash scripts/gpu_run_r2_launches.sh r2
TAG=${1:-r2}
O=gpurun_out
K='regex:score_ldg|score_tma|topk_segments|topk_merge|mean_pool|maxsim_|tc_score|gather_vt5|retrieve_cluster|s2_weights'
python bench.py --steps 20 --warmup 3 --skip-e2e --no-legs --min-replays 3 --min-ms 1 > $O/${TAG}_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 20 --warmup 3 --skip-e2e --no-legs --min-replays 3 --min-ms 1 > $O/${TAG}_ncu_launches.log 2>&1
tail -n 2 $O/${TAG}_ncu_launches.log | cut -c1-300
