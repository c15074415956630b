#!/bin/bash
# K4 stand-alone timing experiment: no next-ray prefetch (64 registers) at 4 CTAs per SM.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1', {k: (round(v['us_per_launch'],1), round(v['hbm_frac'],3)) for k,v in d['kernels'].items() if 'composite' in k})
"; }
sed -i 's/#define COMP_PF(C_) ((C_) <= 8)/#define COMP_PF(C_) (false)/' pcnerf_b200/csrc/composite.cu
python -m pcnerf_b200.build > $O/s2i_build.log 2>&1; tail -1 $O/s2i_build.log
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2i_nopf.json 2>/dev/null; show $O/s2i_nopf.json
sed -i 's/__global__ void __launch_bounds__(256) k_composite_fwd_r(/__global__ void __launch_bounds__(256, 4) k_composite_fwd_r(/' pcnerf_b200/csrc/composite.cu
sed -i 's/__global__ void __launch_bounds__(256) k_composite_bwd_r(/__global__ void __launch_bounds__(256, 4) k_composite_bwd_r(/' pcnerf_b200/csrc/composite.cu
python -m pcnerf_b200.build > $O/s2i_build.log 2>&1; tail -1 $O/s2i_build.log
timeout 300 python scripts/hbm_kernels.py --rays 262144 --S 64 --Ni 128 --flush read > $O/s2i_nopf_occ4.json 2>/dev/null; show $O/s2i_nopf_occ4.json
timeout 300 python scripts/hbm_kernels.py --rays 32768 --S 64 --Ni 128 --flush read > $O/s2i_nopf_occ4_32k.json 2>/dev/null; show $O/s2i_nopf_occ4_32k.json
