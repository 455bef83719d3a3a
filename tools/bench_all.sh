#!/bin/bash
# Runs the per-workload bench lines (device-resident value + roofline; short e2e leg) and collects them.
#   tools/bench_all.sh <outdir> [workloads...]
out=${1:-gpurun_out/bench}; shift
mkdir -p "$out"
wls=${@:-pipeline_imix pipeline_64 nat_steady_64 nat_cold_64 nat_ingress_64 antispoof_64 qos_64 dhcp}
for w in $wls; do
  python bench.py --workload $w --steps ${STEPS:-10} --warmup 3 --no-cpu --no-extra --e2e-steps 1 ${EXTRA} > "$out/$w.json" 2> "$out/$w.err" || echo "FAILED $w"
  python - "$out/$w.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d["roofline"]
    print(f"coop {d.get('nat_ordered_chunks')}", end=" "); print(f'{d["config"]["workload"]:16s} {d["value"]:9.1f} Mpps  {d["ms_per_step"]:.4f} ms/step  top {r["kernel"]} {r["kernel_ms"]} ms frac {r["frac"]}  e2e {(d.get("e2e") or {}).get("value")}  kernels {r["kernels_ms"]}')
except Exception as e:
    print("no result", sys.argv[1], e)
PY
done
