#!/bin/bash
# One B200 call: full GPU suite, smoke, bench lines (round tag $1).  ncu captures: scripts/gpu_run2.sh / gpu_run3.sh.
TAG=${1:-r1h}
O=gpurun_out
python -m pytest tests -q -m gpu > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
python bench.py --impl reference > $O/${TAG}_bench_ref_c2.json 2> $O/${TAG}_ref.err; cut -c1-200 $O/${TAG}_bench_ref_c2.json
python bench.py > $O/${TAG}_bench_c2_n1.json 2> $O/${TAG}_c2.err; tail -2 $O/${TAG}_c2.err; cut -c1-300 $O/${TAG}_bench_c2_n1.json
python bench.py --workload C4 > $O/${TAG}_bench_c4_n1.json 2> $O/${TAG}_c4.err; tail -2 $O/${TAG}_c4.err; cut -c1-200 $O/${TAG}_bench_c4_n1.json
