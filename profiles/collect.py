#!/usr/bin/env python
"""Turns the scratch output of profiles/capture.sh (gpurun_out/<round>_*) into the tracked artefacts:

  profiles/<round>_ncu_<name>.txt    summary of every .ncu-rep (ncu_summary.py)
  profiles/<round>_launches.csv      ncu launch list of the default bench command
  profiles/<round>_clocks.csv        nvidia-smi samples during the plain run
  profiles/<round>_bench*.json       the bench lines themselves
  profiles/ncu_traffic.json          dram__bytes_read+write per launch of each workload's top kernel
                                     (bench.py reports it as roofline.traffic)
  profiles/<round>_results.md        the table of measured numbers

usage: python profiles/collect.py r01        (run here, where ncu can read the reports)
"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


def dram_bytes(rep, kernel_substr):
    h, rows = raw(rep)
    ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out2 = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    units = list(csv.reader(out2.splitlines()))[1]
    for r in rows:
        if kernel_substr in r[ki]:
            return int(float(r[ri].replace(",", "")) * unit[units[ri]] + float(r[wi].replace(",", "")) * unit[units[wi]])
    return None


def main():
    for rep in sorted(glob.glob(os.path.join(SRC, f"{R}_*.ncu-rep"))):
        name = os.path.basename(rep)[len(R) + 1:-len(".ncu-rep")]
        txt = subprocess.run([sys.executable, os.path.join(DST, "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(DST, f"{R}_ncu_{name}.txt"), "w").write(txt)
    for f in glob.glob(os.path.join(SRC, f"{R}_launches.csv")) + glob.glob(os.path.join(SRC, f"{R}_clocks.csv")) + \
            glob.glob(os.path.join(SRC, f"{R}_bench*.json")):
        shutil.copy(f, DST)

    traffic = {}
    for wl, rep, kern, label in (("pipeline_imix", "classify", "k_pipe_classify", "(k_pipe_classify<true, true>)"),
                                 ("dhcp", "dhcp", "k_dhcp_fastpath", "k_dhcp_fastpath"),
                                 ("antispoof_64", "antispoof", "k_antispoof", "k_antispoof")):
        p = os.path.join(SRC, f"{R}_{rep}.ncu-rep")
        if os.path.exists(p):
            b = dram_bytes(p, kern)
            if b:
                traffic[wl] = {label: b}
    json.dump(traffic, open(os.path.join(DST, "ncu_traffic.json"), "w"), indent=1)

    names = ["", "_pipeline_64", "_antispoof_64", "_nat_steady_64", "_nat_cold_64", "_nat_ingress_64", "_qos_64", "_qos_egress_64",
             "_dhcp"]
    lines = [f"# Measured on B200 — round {R}", "",
             "One `python bench.py --workload W` line each (`profiles/%s_bench_W.json`), 2^22 frames per step unless the" % R,
             "workload holds fewer; `value` = device-resident throughput, `e2e` = through the C ABI from a pinned host arena.",
             "Roofline: SURVEY.md §8(d) algorithmic bytes per frame x frames / time, against MEASURED_PEAKS.json (6591.9 GB/s).",
             "",
             "| workload | Mpps (device-resident) | ms/step | roofline frac (whole step) | dominant kernel | its ms | its frac | e2e Mpps |",
             "|---|---|---|---|---|---|---|---|"]
    for n in names:
        p = os.path.join(DST, f"{R}_bench{n}.json")
        if not os.path.exists(p):
            continue
        j = json.loads(open(p).readline())
        r = j["roofline"]
        lines.append("| %s | %.0f | %.4f | %.3f | %s | %.4f | %.3f | %.0f |" % (
            j["config"]["workload"], j["value"], j["ms_per_step"], r["step_frac"], r["kernel"].strip("()"), r["kernel_ms"],
            r["frac"], j["e2e"]["value"]))
    p = os.path.join(DST, f"{R}_bench_reference.json")
    if os.path.exists(p):
        j = json.loads(open(p).readline())
        lines += ["", "Reference arm (`bench.py --impl reference`): %.1f Mpps on %s host threads (%s)." % (
            j["value"], j["cpu_baseline"]["cores"], j["cpu_baseline"]["kind"])]
    p = os.path.join(DST, f"{R}_bench.json")
    if os.path.exists(p):
        j = json.loads(open(p).readline())
        lines += ["", "Per-kernel times of the default workload (CUDA events around every launch, separate pass): `%s`." %
                  json.dumps(j["roofline"]["kernels_ms"]),
                  "DRAM traffic of the dominant kernel (one `ncu --set full` capture): %s bytes per launch = %.0f B/frame." % (
                      j["roofline"]["traffic"], (j["roofline"]["traffic"] or 0) / j["config"]["frames_per_gpu_per_step"]),
                  "Header-split e2e (64-byte header ring, DMA both ways): %s Mpps." % (j.get("e2e_header_split") or {}).get("value")]
    extra = os.path.join(DST, f"{R}_notes.md")
    if os.path.exists(extra):
        lines += ["", open(extra).read()]
    open(os.path.join(DST, f"{R}_results.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
