#!/bin/bash
# A/B of an environment switch on the headline bench line:  tools/ab_env.sh VAR "v1 v2 v1 v2" [bench args]
# prints value / e2e / breakdown per run (run on the GPU box).
var=$1; vals=$2; shift 2
for v in $vals; do
  env "$var=$v" timeout 150 python bench.py "$@" 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', round(d['value'], 2), round(d['e2e']['value'], 2), {k: round(x, 4) for k, x in d['breakdown_ms'].items()})"
done
