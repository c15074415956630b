#!/bin/bash
# usage: gpu_ab.sh TAG:ENV=VAL[,ENV=VAL] ...   -- short same-box A/B of bench.py (training step only)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
for v in "$@"; do
  tag=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-inference --no-fast-mode --no-c4 --no-c5 --no-cpu-baseline > $O/r2_ab_$tag.json 2> $O/r2_ab_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_ab_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], {k: round(v["ms_per_step"],2) for k,v in d.get("kernels",{}).items() if k.startswith("mlp")})
except Exception as e:
    print("$tag failed", e); print(open("$O/r2_ab_$tag.err").read()[-1500:])
PY
done
