#!/bin/bash
# One B200 call: full GPU suite, bench lines, launch list, ncu captures (round tag $1).
TAG=${1:-r1h}
O=gpurun_out
python -m pytest tests -q -m gpu > $O/${TAG}_tests.log 2>&1; tail -4 $O/${TAG}_tests.log
# (compute-sanitizer is closed on this pool: bounds are covered by the parity tests against the oracle)
python bench.py --impl reference > $O/${TAG}_bench_ref_c2.json 2> $O/${TAG}_ref.err; cut -c1-300 $O/${TAG}_bench_ref_c2.json
python bench.py > $O/${TAG}_bench_c2_n1.json 2> $O/${TAG}_c2.err; tail -2 $O/${TAG}_c2.err; cut -c1-400 $O/${TAG}_bench_c2_n1.json
python bench.py --workload C4 > $O/${TAG}_bench_c4_n1.json 2> $O/${TAG}_c4.err; tail -2 $O/${TAG}_c4.err; cut -c1-300 $O/${TAG}_bench_c4_n1.json
python bench.py --steps 20 --warmup 3 --skip-e2e --no-extras > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2.csv \
    python bench.py --steps 20 --warmup 3 --skip-e2e --no-extras > $O/${TAG}_ncu_l.log 2>&1
python scripts/probe_postproc.py > $O/${TAG}_probe_postproc.json 2> $O/${TAG}_pp.err && \
ncu --set full --clock-control none --import-source on -k regex:"rerank_order|page_vote|layout_assign" -c 5 -f -o $O/${TAG}_postproc \
    python scripts/probe_postproc.py > $O/${TAG}_ncu_pp.log 2>&1
python scripts/probe_pix2struct.py > $O/${TAG}_probe_p2s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"p2s_" -c 4 -f -o $O/${TAG}_pix2struct \
    python scripts/probe_pix2struct.py > $O/${TAG}_ncu_p2s.log 2>&1
ls -la $O/${TAG}_*
