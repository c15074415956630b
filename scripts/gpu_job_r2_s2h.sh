#!/bin/bash
# K4 / K2' stand-alone timing: L2 flush by read vs write, the alternative K4 lane split, and K4 forward at 4 CTAs per SM.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1', {k: (round(v['us_per_launch'],1), round(v['hbm_frac'],3)) for k,v in d['kernels'].items()})
"; }
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush write > $O/s2h_write.json 2>/dev/null; show $O/s2h_write.json
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2h_read.json 2>/dev/null; show $O/s2h_read.json
PCNERF_K4_SHAPE=1 timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2h_read_alt.json 2>/dev/null; show $O/s2h_read_alt.json
sed -i 's/__global__ void __launch_bounds__(256) k_composite_fwd_r(/__global__ void __launch_bounds__(256, 4) k_composite_fwd_r(/' pcnerf_b200/csrc/composite.cu
python -m pcnerf_b200.build > $O/s2h_build.log 2>&1; tail -1 $O/s2h_build.log
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2h_read_occ4.json 2>/dev/null; show $O/s2h_read_occ4.json
PCNERF_K4_SHAPE=1 timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2h_read_occ4_alt.json 2>/dev/null; show $O/s2h_read_occ4_alt.json
