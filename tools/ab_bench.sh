#!/bin/bash
# A/B builds of the library on the same GPU, alternating (power capping makes single runs hard to compare):
#   ROUNDS=2 tools/ab_bench.sh nphm_b200/libnphm_b200_A.so nphm_b200/libnphm_b200.so ...
R=${ROUNDS:-2}
for r in $(seq $R); do
  for lib in "$@"; do
    NPHM_B200_LIB=$PWD/$lib timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'sdf_ms %.1f  step %.1f  sm %.0f MHz  P %.0f W  cycles %.1f M  pruned %.1f ms' % (d['sdf_ms'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['sdf_ms']*d['clocks']['sm_mhz']/1e3, d['pruned_opt_in']['sdf_ms']), d['clocks']['reasons'])"
  done
done
