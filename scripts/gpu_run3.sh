#!/bin/bash
TAG=${1:-r1h}
O=gpurun_out
python -m pytest tests/test_visual_pack_gpu.py -q -m gpu > $O/${TAG}_vis_tests.log 2>&1; tail -2 $O/${TAG}_vis_tests.log
python scripts/probe_visual.py > $O/${TAG}_vis.log 2>&1; tail -1 $O/${TAG}_vis.log
ncu --set full --clock-control none --import-source on -k regex:"visual_resize" --launch-skip 6 -c 2 -f -o $O/${TAG}_visual_h \
    python scripts/probe_visual.py > $O/${TAG}_ncu_v.log 2>&1
