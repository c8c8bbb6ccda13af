#!/bin/bash
# A/B harness: time the bench headline for several builds of libmcgp (MCGP_LIB_PATH override)
for lib in "$@"; do
  MCGP_LIB_PATH=$PWD/monte-carlo-gp_b200/$lib python bench.py --no-cpu-baseline --steps 3 --sims-per-step 6000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']/1e6,2), 'M races/s frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
