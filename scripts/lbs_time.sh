#!/bin/bash
# usage: scripts/lbs_time.sh "PAIR DEBUG STAGED" ...   (A/B timing of the LBS forward; each run is killed after 40 s)
for v in "$@"; do set -- $v
  out=$(DPB_LBS_PAIR=$1 DPB_LBS_DEBUG=$2 DPB_LBS_STAGED=${3:-1} timeout -s KILL 40 python bench.py --workload lbs --no-cpu --steps 3 2>&1 | tail -1)
  python - "$v" "$out" <<'PY'
import sys, json
try:
    d = json.loads(sys.argv[2]); print('pair/debug/staged =', sys.argv[1], '->', round(d['roofline']['ms'], 4), 'ms')
except Exception:
    print('pair/debug/staged =', sys.argv[1], '-> HANG or error:', sys.argv[2][-200:])
PY
done
