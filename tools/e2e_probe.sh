#!/bin/bash
# What bounds the end-to-end (host arena -> GPU -> host arena) leg on this box?
#   gpurun --timeout 900 -- 'bash tools/e2e_probe.sh'
set -u
mkdir -p gpurun_out
{
    echo "== thp: $(cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null)  defrag: $(cat /sys/kernel/mm/transparent_hugepage/defrag 2>/dev/null)"
    echo "== iommu groups: $(ls /sys/kernel/iommu_groups 2>/dev/null | wc -l)  cmdline: $(cat /proc/cmdline 2>/dev/null | tr ' ' '\n' | grep -i -E 'iommu|hugepage' | tr '\n' ' ')"
    lscpu | grep -E "NUMA|Model name|^CPU\(s\)"
    nvidia-smi topo -m 2>/dev/null | head -12
    nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
    grep -E "AnonHugePages|HugePages_Total|Hugepagesize" /proc/meminfo
} > gpurun_out/e2e_probe_sys.txt 2>&1
cat gpurun_out/e2e_probe_sys.txt
for arena in pinned thp; do
    for w in pipeline_imix pipeline_64; do
        python bench.py --workload $w --steps 5 --e2e-steps 4 --no-cpu --arena $arena 2> gpurun_out/e2e_probe_$arena.$w.err |
            python -c "import sys,json; j=json.loads(sys.stdin.readline()); print('$arena', '$w', 'dev', j['value'], 'e2e', j['e2e']['value'], 'hs', (j.get('e2e_header_split') or {}).get('value'))"
        grep -E "AnonHugePages" /proc/meminfo
    done
done
