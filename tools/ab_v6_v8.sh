export NPHM_BENCH_NO_SAMPLER=
for k in v6 v8 v6 v8; do
  NPHM_TC_KERNEL=$k timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$k', 'sdf_ms', round(d['sdf_ms'],2), 'frac', round(d['roofline']['frac'],4), 'clk', d['clocks']['sm_mhz'], 'pruned', round(d['pruned_opt_in']['sdf_ms'],2), 'maxdiff', d['pruned_opt_in']['max_abs_diff_vs_dense'])
"
done
