#!/bin/bash
# Round-2 ncu evidence (one gpurun call):  gpurun --timeout 1500 -- 'bash scripts/gpu_run_r2_ncu.sh r2'
#  1. launch list of the default bench step (device-resident steps only)   -> ${TAG}_launches_c2.csv
#  2. ncu --set full of every kernel scripts/ncu_targets_r2.py launches     -> ${TAG}_kernels.ncu-rep   (cache-control all = cold)
#  3. the C2 / C3 streaming kernel again with --cache-control none          -> ${TAG}_score_warm.ncu-rep
TAG=${1:-r2}
O=gpurun_out
K='regex:score_ldg|score_tma|topk_segments|topk_merge|mean_pool|maxsim_|tc_score|gather_vt5|retrieve_cluster|s2_weights'
python scripts/ncu_targets_r2.py all > $O/${TAG}_targets_plain.log 2>&1 && \
ncu --set full --clock-control none -k "$K" -o /tmp/${TAG}_kernels python scripts/ncu_targets_r2.py all > $O/${TAG}_ncu_kernels.log 2>&1
# gpurun brings back at most 64 MiB: the capture stays on the box, its raw page (every metric + the stall samples) travels
ncu -i /tmp/${TAG}_kernels.ncu-rep --page raw --csv > $O/${TAG}_kernels_raw.csv 2>/dev/null
tail -n 2 $O/${TAG}_ncu_kernels.log
python scripts/ncu_targets_r2.py score > $O/${TAG}_targets_score_plain.log 2>&1 && \
ncu --set full --clock-control none --cache-control none -k 'regex:score_ldg' -o /tmp/${TAG}_score_warm python scripts/ncu_targets_r2.py score > $O/${TAG}_ncu_score_warm.log 2>&1
ncu -i /tmp/${TAG}_score_warm.ncu-rep --page raw --csv > $O/${TAG}_score_warm_raw.csv 2>/dev/null
tail -n 2 $O/${TAG}_ncu_score_warm.log
python bench.py --steps 20 --warmup 3 --skip-e2e --no-legs --min-replays 3 --min-ms 1 > $O/${TAG}_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 20 --warmup 3 --skip-e2e --no-legs --min-replays 3 --min-ms 1 > $O/${TAG}_ncu_launches.log 2>&1
tail -n 2 $O/${TAG}_ncu_launches.log
ls -la $O/${TAG}_*; du -sh $O
