#!/bin/bash
# ncu --set full of the shipped data-gradient (k_tc_rowgemm2<DGRAD2>) and weight-gradient (k_tc_wgrad2) kernels: DRAM traffic per launch
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:"^k_tc_rowgemm2$" -s 300 -c 4 -o /tmp/s2p_ncu -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-inference --no-fast-mode --no-c4 --no-c5 --graph off > $O/s2p_ncu.log 2>&1
tail -1 $O/s2p_ncu.log | cut -c1-160
ncu -i /tmp/s2p_ncu.ncu-rep --page raw --csv > $O/s2p_ncu_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/s2p_ncu_raw.csv")))
h = rows[0]; ix = {c: i for i, c in enumerate(h)}
for r in rows[2:]:
    print(r[ix["Kernel Name"]][:44], r[ix["gpu__time_duration.sum"]], "rd", r[ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_read.sum"]], "wr", r[ix["dram__bytes_write.sum"]], rows[1][ix["dram__bytes_write.sum"]])
PY
