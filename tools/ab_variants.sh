#!/bin/bash
# A/B of experiment builds of the library on the three-height call: tools/ab_variants.sh [lib.so ...]
# (the default library first).  One line per library: ms of three launches, then of the one-pass kernel per tile / chunk.
show() {
  python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['lib'], round(d['three_launches_ms'], 3), {k[9:]: round(v['ms'], 3) for k, v in d.items() if isinstance(v, dict)})"
}
python tools/variants_probe.py 2>/dev/null | show
for l in "$@"; do MOMLEVEL_B200_LIB=$l python tools/variants_probe.py 2>/dev/null | show; done
