#!/bin/bash
# A/B of library builds on the headline step (bench.py, kernel-only): tools/ab_headline.sh lib.so [rounds]
# alternates default / lib for `rounds` rounds; prints ms per step of the 20-step timing and of the 200-step self-check
lib=$1; rounds=${2:-3}
one() {
  python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'], 4), round(d['selfcheck_200_steps']['ms_per_step'], 4), d['selfcheck_200_steps']['clocks'].get('sm_mhz'), d['selfcheck_200_steps']['clocks'].get('reasons'))"
}
for i in $(seq $rounds); do
  one default
  MOMLEVEL_B200_LIB=$lib one $lib
done
