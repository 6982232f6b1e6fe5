#!/bin/bash
# usage: tools/ab_env.sh "VAR=1 VAR2=x" "VAR=2" ...   -- one bench.py run per environment setting
i=0
for e in "$@"; do
  i=$((i+1))
  env $e python bench.py --steps 3 --warmup 3 --no-qq --no-prob 2>/dev/null > gpurun_out/abe_$i.json || echo "$e failed"
  python - "$e" $i <<PY
import json, sys
d = json.load(open("gpurun_out/abe_%s.json" % sys.argv[2]))
print("%-40s step %.1f ms  hot %.1f  %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"],
      {k: round(v, 1) for k, v in d.get("phases_ms", {}).items()}))
PY
done
