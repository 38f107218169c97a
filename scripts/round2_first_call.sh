#!/bin/bash
# First B200 call of the next round: what the round-1 GPU budget did not cover (open items).
#   gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh r2a'
TAG=${1:-r2a}
O=gpurun_out
timeout 150 python -m pytest tests -q -m gpu -x > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
# C5 with the recall@k line (bench.py run_corpus: recall_at_k_vs_fp32), C1 / C2 with the final round-1 code
timeout 150 python bench.py --workload C5 > $O/${TAG}_bench_c5_n1.json 2> $O/${TAG}_c5.err; tail -1 $O/${TAG}_c5.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench_c5_n1.json").read().strip().splitlines()[-1])
print("C5", d["value"], d.get("recall_at_k_vs_fp32"))
PY
timeout 100 python bench.py --workload C1 > $O/${TAG}_bench_c1_n1.json 2> $O/${TAG}_c1.err; cut -c1-200 $O/${TAG}_bench_c1_n1.json
timeout 100 python bench.py > $O/${TAG}_bench_c2_n1.json 2> $O/${TAG}_c2.err; cut -c1-200 $O/${TAG}_bench_c2_n1.json
# memory and race checks of the kernels on the small parity cases (never run in round 1)
for TOOL in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest -q -m gpu -x \
    tests/test_score_topk_gpu.py tests/test_retriever_gpu.py tests/test_postproc_gpu.py tests/test_s2chunker_gpu.py tests/test_chunker_gpu.py \
    -k "golden or small or special or empty" > $O/${TAG}_sanitizer_${TOOL}.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/${TAG}_sanitizer_${TOOL}.log | tail -3
done
