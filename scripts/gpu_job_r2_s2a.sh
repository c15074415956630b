#!/bin/bash
# Round-2 (second session) check job, 1 GPU: GPU test suite, default bench, launch list + per-class times of the closed-form step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x > $O/s2a_gputest.log 2>&1; tail -4 $O/s2a_gputest.log
echo "== bench default"; timeout 900 python bench.py > $O/s2a_bench_default.json 2> $O/s2a_bench_default.err; tail -3 $O/s2a_bench_default.err; cut -c1-400 $O/s2a_bench_default.json
echo "== bench affine"; timeout 600 python bench.py --precision affine --no-c4 --no-c5 --no-cpu-baseline > $O/s2a_bench_affine.json 2> $O/s2a_bench_affine.err; tail -3 $O/s2a_bench_affine.err; cut -c1-300 $O/s2a_bench_affine.json
echo "== ncu launch list (closed-form step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/s2a_launches_affine.csv python bench.py --precision affine --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-c4 --no-c5 --graph off > $O/s2a_ncu_launch.log 2>&1
tail -1 $O/s2a_ncu_launch.log | cut -c1-200
echo "== ncu full (closed-form kernels)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_aff" -c 60 -o $O/s2a_ncu_aff -f python scripts/run_affine_once.py > $O/s2a_ncu_full.log 2>&1
tail -1 $O/s2a_ncu_full.log | cut -c1-200
ls -la $O
