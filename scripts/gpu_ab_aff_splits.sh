cd "${GRAFT_REPO_ROOT:-/root/repo}"
for sp in 2 6 3; do
PCNERF_AFF_SPLITS=$sp timeout 300 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline --no-inference > gpurun_out/s2f_$sp.json 2>/dev/null
python -c "
import json
d = json.loads(open('gpurun_out/s2f_$sp.json').read().strip().splitlines()[-1])
print('splits $sp', d['ms_per_step'], {k: round(v['ms_per_step'], 4) for k, v in d['kernels'].items() if 'aff' in k})
"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_aff_ -c 400 --csv --log-file gpurun_out/s2f_launches.csv python bench.py --precision affine --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-c4 --no-c5 --graph off > /dev/null 2>&1
