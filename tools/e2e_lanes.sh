#!/bin/bash
# Usage (under gpurun): bash tools/e2e_lanes.sh -- e2e ms/step with one / two extraction lanes, with / without dsx_survey
for L in 1 2; do
  for S in "" "--no-survey-call"; do
    DSX_H2D_LANES=$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-bruteforce --h2d-chunk ${CHUNK:-4} $S > /tmp/e2e.json 2>/tmp/e2e.err
    python - "$L" "$S" <<'P'
import json, sys
try:
    d = json.load(open('/tmp/e2e.json'))
    print("lanes %s %-18s resident %.2f ms  e2e %.2f ms" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("lanes", sys.argv[1], sys.argv[2], "failed", e, open('/tmp/e2e.err').read()[-600:])
P
  done
done
