#!/bin/bash
# Runs bench.py once per library variant (built by tools/build_variants.sh) and prints the step / hot-kernel times.
# usage (on the GPU box): tools/ab_variants.sh base occ3 u2 ...     ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset F4LA_B200_LIB; else export F4LA_B200_LIB=$PWD/msolve_b200/csrc/variants/libf4la_b200_$v.so; fi
  python bench.py --steps 3 --warmup 3 --no-qq --no-prob 2>/dev/null > gpurun_out/ab_$v.json || echo "$v failed"
  python - "$v" <<PY
import json, sys
d = json.load(open("gpurun_out/ab_%s.json" % sys.argv[1]))
print("%-10s step %.1f ms  hot kernel %.1f ms  phases %s" % (sys.argv[1], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"],
      {k: round(v, 1) for k, v in d.get("phases_ms", {}).items()}))
PY
done
