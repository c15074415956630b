#!/bin/bash
# ncu --set full of the stand-alone HBM kernels (K2', K4 forward / backward) at 262,144 rays x (64 + 128); CSV exports only.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 > $O/s2g_hbm_262144.json 2>/dev/null; cut -c1-1500 $O/s2g_hbm_262144.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_composite_fwd_r|k_composite_bwd_r|k_sample_encode_fine" -s 6 -c 8 -o /tmp/s2g_ncu -f python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --reps 1 > $O/s2g_ncu.log 2>&1
tail -2 $O/s2g_ncu.log | cut -c1-200
ncu -i /tmp/s2g_ncu.ncu-rep --page raw --csv > $O/s2g_ncu_raw.csv 2>/dev/null
ncu -i /tmp/s2g_ncu.ncu-rep --page source --csv --kernel-name regex:k_composite_fwd_r > $O/s2g_ncu_src_k4fwd.csv 2>/dev/null
ncu -i /tmp/s2g_ncu.ncu-rep --page source --csv --kernel-name regex:k_composite_bwd_r > $O/s2g_ncu_src_k4bwd.csv 2>/dev/null
ncu -i /tmp/s2g_ncu.ncu-rep --page source --csv --kernel-name regex:k_sample_encode_fine > $O/s2g_ncu_src_k2fine.csv 2>/dev/null
du -sh $O; ls -la $O | tail -8
