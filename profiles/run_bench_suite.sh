#!/bin/bash
# Regenerates profiles/r01_bench/*.json on a 1-GPU box:  gpurun -- 'bash profiles/run_bench_suite.sh'
# (results land in gpurun_out/bench/, copy them to profiles/r01_bench/ afterwards)
mkdir -p gpurun_out/bench
run() { out=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/bench/$out.err | tail -1 > gpurun_out/bench/$out.json; }
run c2_nich --workload c2_nich --steps 20
run c2_reference_arm --impl reference --workload c2_nich --steps 3 --warmup 1
run c2_nich_materialised --workload c2_nich --materialise --steps 10 --no-cpu
run c2_nich_sweep --workload c2_nich --sweep --steps 20 --no-cpu
run c1_dd --workload c1_dd --steps 20
run c1_dd_steady --workload c1_dd_steady --steps 10
run c3_crosscat --workload c3_crosscat --steps 10
run c3_crosscat_sweep --workload c3_crosscat --sweep --steps 10 --no-cpu
run c4_dpd --workload c4_dpd --steps 10
run c4_dpd_sweep --workload c4_dpd --sweep --steps 10 --no-cpu
run c5_niw --workload c5_niw --steps 10
for f in gpurun_out/bench/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[1].split('/')[-1], round(d.get('ms_per_step',0),4), '%.3e'%d.get('value',0), 'e2e %.3e'%((d.get('e2e') or {}).get('value') or 0), 'cpu %.3e'%((d.get('cpu_baseline') or {}).get('value') or 0))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
