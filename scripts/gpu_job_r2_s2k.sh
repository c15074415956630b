#!/bin/bash
# closed-form engine: second moments on mma.sync 3xTF32 vs packed FFMA2 -- parity tests and A/B bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== tests (tc moments)"; timeout 600 python -m pytest tests/test_gpu_affine.py tests/test_gpu_c1.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/s2k_gputest.log 2>&1; tail -15 $O/s2k_gputest.log
for m in tc ffma; do
PCNERF_AFF_MOMENTS=$m timeout 300 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline --no-inference > $O/s2k_$m.json 2>/dev/null
python -c "
import json
d = json.loads(open('gpurun_out/s2k_$m.json').read().strip().splitlines()[-1])
print('moments $m', d['ms_per_step'], {k: round(v['ms_per_step'], 4) for k, v in d['kernels'].items() if 'aff' in k})
"
done
