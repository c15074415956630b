#!/bin/bash
# Round-2 (second session) job B, 1 GPU: full GPU test suite, closed-form bench with per-class times, launch list and
# ncu --set full of the closed-form kernels (exported to CSV on the box: gpurun_out/ must stay below 64 MiB).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q > $O/s2b_gputest.log 2>&1; tail -8 $O/s2b_gputest.log
echo "== bench affine"; timeout 600 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline > $O/s2b_bench_affine.json 2> $O/s2b_bench_affine.err; tail -3 $O/s2b_bench_affine.err; cut -c1-300 $O/s2b_bench_affine.json
echo "== bench tc fast-mode block only"; timeout 600 python bench.py --no-c4 --no-c5 --no-cpu-baseline --no-inference > $O/s2b_bench_tc_fast.json 2> $O/s2b_bench_tc_fast.err; tail -3 $O/s2b_bench_tc_fast.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/s2b_bench_tc_fast.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("fast_mode"))[:3000])
PY
echo "== ncu launch list (closed-form step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/s2b_launches_affine.csv python bench.py --precision affine --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-c4 --no-c5 --graph off > $O/s2b_ncu_launch.log 2>&1
tail -1 $O/s2b_ncu_launch.log | cut -c1-200
echo "== ncu full (closed-form kernels)"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_aff" -c 30 -o /tmp/s2b_ncu_aff -f python scripts/run_affine_once.py > $O/s2b_ncu_full.log 2>&1
tail -1 $O/s2b_ncu_full.log | cut -c1-200
ncu -i /tmp/s2b_ncu_aff.ncu-rep --page raw --csv > $O/s2b_ncu_aff_raw.csv 2>/dev/null
for k in k_affine_moments_rays k_affine_grad_rays k_affine_apply_rays; do
  ncu -i /tmp/s2b_ncu_aff.ncu-rep --page source --csv --kernel-name $k > $O/s2b_ncu_src_$k.csv 2>/dev/null
done
du -sh $O; ls -la $O
