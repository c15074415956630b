#!/bin/bash
# Round-2 (second session) job D: fused per-layer kernels of the closed-form engine -- parity tests, A/B bench against the
# unfused chain (PCNERF_AFF_FUSED=0), launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== affine tests"; timeout 900 python -m pytest tests/test_gpu_affine.py tests/test_gpu_c1.py tests/test_gpu_fullsize.py -m gpu -q > $O/s2d_gputest.log 2>&1; tail -8 $O/s2d_gputest.log
echo "== bench affine fused"; timeout 600 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline > $O/s2d_bench_affine_fused.json 2> $O/s2d_bench_affine_fused.err; tail -3 $O/s2d_bench_affine_fused.err; cut -c1-200 $O/s2d_bench_affine_fused.json
echo "== bench affine unfused"; PCNERF_AFF_FUSED=0 timeout 600 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline > $O/s2d_bench_affine_unfused.json 2> $O/s2d_bench_affine_unfused.err; tail -3 $O/s2d_bench_affine_unfused.err; cut -c1-200 $O/s2d_bench_affine_unfused.json
python - <<'PY'
import json
for n in ("fused", "unfused"):
    d = json.loads(open("gpurun_out/s2d_bench_affine_%s.json" % n).read().strip().splitlines()[-1])
    print(n, d["ms_per_step"], {k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()}, d["inference"]["ms_per_frame"])
PY
echo "== ncu launch list (closed-form step, fused)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/s2d_launches_affine.csv python bench.py --precision affine --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-c4 --no-c5 --graph off > $O/s2d_ncu_launch.log 2>&1
tail -1 $O/s2d_ncu_launch.log | cut -c1-200
du -sh $O
