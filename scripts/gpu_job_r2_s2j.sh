#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sample_encode_fine" -c 2 -o /tmp/s2j_ncu -f python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --reps 1 > $O/s2j_ncu.log 2>&1
tail -2 $O/s2j_ncu.log | cut -c1-200
ncu -i /tmp/s2j_ncu.ncu-rep --page raw --csv > $O/s2j_ncu_raw.csv 2>/dev/null
ncu -i /tmp/s2j_ncu.ncu-rep --page source --csv > $O/s2j_ncu_src_k2fine.csv 2>/dev/null
ncu -i /tmp/s2j_ncu.ncu-rep --page details > $O/s2j_ncu_details.txt 2>/dev/null
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2j_hbm.json 2>/dev/null; cut -c1-1200 $O/s2j_hbm.json
ls -la $O | tail -6
