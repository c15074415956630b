#!/bin/bash
# timing experiment: closed-form moment kernels with free producers (wrong results) -- what the consumers alone would take
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for m in ffma tc; do for d in 0 1; do
PCNERF_AFF_DEBUG=$d PCNERF_AFF_MOMENTS=$m timeout 300 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline --no-inference --steps 3 > gpurun_out/s2n_${m}_$d.json 2>/dev/null
python -c "
import json
d = json.loads(open('gpurun_out/s2n_${m}_$d.json').read().strip().splitlines()[-1])
print('moments $m debug $d', round(d['ms_per_step'],3), {k: round(v['ms_per_step'], 4) for k, v in d['kernels'].items() if 'aff' in k})
"
done; done
