#!/bin/bash
# A/B of experiment builds on the elementwise kernels: tools/ab_stream.sh lib.so ...  (default library first)
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[0])
print('$1', {k: v['hbm_frac'] for k, v in d.items() if isinstance(v, dict)})"; }
python tools/spice_probe.py 2>/dev/null | show default
for l in "$@"; do MOMLEVEL_B200_LIB=$l python tools/spice_probe.py 2>/dev/null | show $l; done
