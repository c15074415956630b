#!/bin/bash
# Round-2 final measurement job (1 GPU): full GPU test suite, default bench, error study, stand-alone HBM kernels, ncu evidence.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q > $O/r2f_gputest.log 2>&1; tail -6 $O/r2f_gputest.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; timeout 900 python bench.py > $O/r2f_bench_default.json 2> $O/r2f_bench_default.err; tail -3 $O/r2f_bench_default.err; cut -c1-300 $O/r2f_bench_default.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_reference.json 2>&1; cut -c1-300 $O/r2f_bench_reference.json
echo "== tc error"; timeout 600 python scripts/tc_error_c2.py > $O/r2f_tc_error.json 2> $O/r2f_tc_error.err; cut -c1-700 $O/r2f_tc_error.json
echo "== hbm kernels"; timeout 600 python scripts/hbm_kernels.py > $O/r2f_hbm_kernels.json 2>/dev/null; timeout 600 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 > $O/r2f_hbm_kernels_262144rays.json 2>/dev/null; tail -c 1200 $O/r2f_hbm_kernels_262144rays.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-fast-mode --no-c4 --no-c5 --graph off > $O/r2f_ncu_launch.log 2>&1
tail -1 $O/r2f_ncu_launch.log | cut -c1-200
echo "== ncu full (pair forward, DGRAD2, wgrad)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_rowgemm2|k_tc_wgrad$" -s 60 -c 8 -o $O/r2f_ncu_gemms -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-fast-mode --no-c4 --no-c5 --graph off > $O/r2f_ncu_full.log 2>&1
tail -1 $O/r2f_ncu_full.log | cut -c1-200
