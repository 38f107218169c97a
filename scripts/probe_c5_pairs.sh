for rows in 1250000 2500000 5000000 10000000; do
  for ctas in 1 2; do
    RDV_TC_CTAS=$ctas timeout 300 python bench.py --workload C5 --corpus-rows $rows --steps 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rows', $rows, 'ctas', $ctas, 'q/s %.0f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'], d['clocks'])
"
  done
done
