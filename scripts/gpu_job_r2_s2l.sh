#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_affine_moments_rays" -c 2 -o /tmp/s2l_ncu -f python scripts/run_affine_once.py > $O/s2l_ncu.log 2>&1
tail -2 $O/s2l_ncu.log | cut -c1-200
ncu -i /tmp/s2l_ncu.ncu-rep --page details > $O/s2l_ncu_details.txt 2>/dev/null
ncu -i /tmp/s2l_ncu.ncu-rep --page source --csv > $O/s2l_ncu_src.csv 2>/dev/null
ncu -i /tmp/s2l_ncu.ncu-rep --page raw --csv > $O/s2l_ncu_raw.csv 2>/dev/null
ls -la $O | tail -5
