#!/bin/bash
# Round-2 closing measurement job (1 GPU): full GPU test suite, smoke, default bench (+ reference arm), stand-alone HBM kernels,
# launch lists of the tensor-core step and of the closed-form step, ncu --set full of the closed-form kernels (CSV exports only:
# gpurun_out/ must stay below 64 MiB).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q > $O/f2_gputest.log 2>&1; tail -5 $O/f2_gputest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; timeout 1200 python bench.py > $O/f2_bench_default.json 2> $O/f2_bench_default.err; tail -3 $O/f2_bench_default.err; cut -c1-300 $O/f2_bench_default.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/f2_bench_reference.json 2>&1; cut -c1-300 $O/f2_bench_reference.json
echo "== hbm kernels"; timeout 600 python scripts/hbm_kernels.py > $O/f2_hbm_kernels.json 2>/dev/null; timeout 600 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 > $O/f2_hbm_kernels_262144rays.json 2>/dev/null; timeout 600 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush write > $O/f2_hbm_kernels_262144rays_writeflush.json 2>/dev/null; tail -c 1000 $O/f2_hbm_kernels_262144rays.json
echo "== ncu launch list (tensor-core step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/f2_launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-fast-mode --no-c4 --no-c5 --graph off > $O/f2_ncu_launch_tc.log 2>&1
tail -1 $O/f2_ncu_launch_tc.log | cut -c1-160
echo "== ncu launch list (closed-form step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/f2_launches_closed_form.csv python bench.py --precision affine --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-c4 --no-c5 --graph off > $O/f2_ncu_launch_cf.log 2>&1
tail -1 $O/f2_ncu_launch_cf.log | cut -c1-160
echo "== ncu full (closed-form kernels)"
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_aff" -c 40 -o /tmp/f2_ncu_aff -f python scripts/run_affine_once.py > $O/f2_ncu_full.log 2>&1
tail -1 $O/f2_ncu_full.log | cut -c1-160
ncu -i /tmp/f2_ncu_aff.ncu-rep --page raw --csv > $O/f2_ncu_closed_form_raw.csv 2>/dev/null
du -sh $O
