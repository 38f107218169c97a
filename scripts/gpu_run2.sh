#!/bin/bash
# visual pack: both horizontal kernels (tests + timing), launch list of the C2 step, ncu of the newest kernels (tag $1)
TAG=${1:-r1h}
O=gpurun_out
python -m pytest tests/test_visual_pack_gpu.py tests/test_retriever_gpu.py -q -m gpu > $O/${TAG}_vis4_tests.log 2>&1; tail -2 $O/${TAG}_vis4_tests.log
RDV_VISUAL_H=3 python -m pytest tests/test_visual_pack_gpu.py -q -m gpu > $O/${TAG}_vis3_tests.log 2>&1; tail -2 $O/${TAG}_vis3_tests.log
python scripts/probe_visual.py > $O/${TAG}_vis4.log 2>&1; tail -1 $O/${TAG}_vis4.log
RDV_VISUAL_H=3 python scripts/probe_visual.py > $O/${TAG}_vis3.log 2>&1; tail -1 $O/${TAG}_vis3.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_|gather_vt5|topk_segments" -c 400 --csv \
    --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 20 --warmup 3 --skip-e2e --no-extras > $O/${TAG}_ncu_l.log 2>&1
python scripts/ncu_targets.py > $O/${TAG}_targets.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rerank_order|page_vote|layout_assign" -c 4 -f -o $O/${TAG}_postproc \
    python scripts/ncu_targets.py > $O/${TAG}_ncu_pp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"visual_resize_h" --launch-skip 3 -c 1 -f -o $O/${TAG}_visual_h4 \
    python scripts/probe_visual.py > $O/${TAG}_ncu_v4.log 2>&1
ls -la $O/${TAG}_* | tail -20
