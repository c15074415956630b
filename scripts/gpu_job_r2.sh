#!/bin/bash
# One gpurun call's worth of round-2 measurements (1 GPU).  Everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q > $O/r2_gputest_full.log 2>&1; tail -25 $O/r2_gputest_full.log
echo "== bench default"; timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_bench_default.json 2> $O/r2_bench_default.err; tail -3 $O/r2_bench_default.err; cut -c1-1500 $O/r2_bench_default.json
for v in "PAIRS0:PCNERF_TC_PAIRS=0" "PAIRS1:PCNERF_TC_PAIRS=1" "CORR0:PCNERF_TC_CORRECT=0" "LANES3:PCNERF_TC_LANES=3" "LANES4:PCNERF_TC_LANES=4" "PAIRS1LANES3:PCNERF_TC_PAIRS=1 PCNERF_TC_LANES=3"; do
  tag=${v%%:*}; envs=${v#*:}
  echo "== bench A/B $tag"
  env $envs timeout 600 python bench.py --steps 5 --warmup 3 --no-inference --no-fast-mode --no-c4 --no-c5 --no-cpu-baseline > $O/r2_bench_ab_$tag.json 2> $O/r2_bench_ab_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_bench_ab_$tag.json"))
    print("$tag", "ms/step", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], {k: round(v["ms_per_step"],2) for k,v in d.get("kernels",{}).items()})
except Exception as e:
    print("$tag failed", e)
PY
done
echo "== k1 scale"; timeout 600 python scripts/time_k1_scale.py > $O/r2_k1_scale.json 2> $O/r2_k1_scale.err; cat $O/r2_k1_scale.json; tail -2 $O/r2_k1_scale.err
echo "== tc error"; timeout 600 python scripts/tc_error_c2.py > $O/r2_tc_error.json 2> $O/r2_tc_error.err; cat $O/r2_tc_error.json
echo "== hbm kernels"; timeout 600 python scripts/hbm_kernels.py > $O/r2_hbm_kernels.json 2> $O/r2_hbm_kernels.err; tail -c 1500 $O/r2_hbm_kernels.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-fast-mode --no-c4 --no-c5 --graph off > $O/r2_ncu_launch.log 2>&1
tail -2 $O/r2_ncu_launch.log | cut -c1-300
echo "== ncu full (forward row GEMM, fold kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_rowgemm|k_tc_fold" -s 40 -c 6 -o $O/r2_ncu_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-fast-mode --no-c4 --no-c5 --graph off > $O/r2_ncu_full.log 2>&1
tail -2 $O/r2_ncu_full.log | cut -c1-300
ls -la $O | tail -15
