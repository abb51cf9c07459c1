#!/bin/bash
# Usage (under gpurun): bash tools/e2e_sweep.sh  -- e2e ms/step of the 64-image survey for several H2D chunk sizes, with and
# without the single dsx_survey call
for C in ${CHUNKS:-2 4 8 16}; do
  for S in "" "--no-survey-call"; do
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-bruteforce --h2d-chunk $C $S > /tmp/e2e.json 2>/tmp/e2e.err
    python - "$C" "$S" <<'P'
import json, sys
try:
    d = json.load(open('/tmp/e2e.json'))
    print("chunk %s %-18s resident %.2f ms  e2e %.2f ms" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("chunk", sys.argv[1], sys.argv[2], "failed", e, open('/tmp/e2e.err').read()[-400:])
P
  done
done
