"""Reads `ncu --set full` reports (gpurun_out/*.ncu-rep or profiles/*.ncu-rep) with `ncu -i ... --page raw --csv` and writes the
per-launch numbers the roofline discussion quotes into profiles/ncu_summary.json (bench.py reads that file for
`roofline.traffic`; it never profiles anything itself).
usage: python tools/ncu_summarize.py TAG=path.ncu-rep ... [--dominant TAG:index] [--out profiles/ncu_summary.json] [--merge]
(--merge keeps the tags already in the output file)"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {
    "duration_us": ("gpu__time_duration.sum", None),
    "dram_read_bytes": ("dram__bytes_read.sum", None),
    "dram_write_bytes": ("dram__bytes_write.sum", None),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
    "tensor_pipe_active_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", None),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", None),
    "l2_to_sm_bytes": ("l1tex__m_xbar2l1tex_read_bytes.sum", None),
    "l2_hit_rate_pct": ("lts__t_sector_hit_rate.pct", None),
    "registers_per_thread": ("launch__registers_per_thread", None),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", None),
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
         "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6, "%": 1.0, "": 1.0, "register/thread": 1.0}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in body:
        d = dict(kernel=r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", ""),
                 grid=r[col["Grid Size"]], block=r[col["Block Size"]])
        for k, (m, _) in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                u = units[col[m]]
                d[k] = float(r[col[m]].replace(",", "")) * SCALE.get(u, 1.0)
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
            if d.get("duration_us"):
                d["dram_gbs"] = d["dram_bytes"] / d["duration_us"] / 1e3
        launches.append(d)
    return launches


def main():
    args = [a for a in sys.argv[1:]]
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_summary.json")
    dominant = None
    if "--out" in args:
        i = args.index("--out"); out_path = args[i + 1]; del args[i:i + 2]
    if "--dominant" in args:
        i = args.index("--dominant"); dominant = args[i + 1]; del args[i:i + 2]
    kernels = {}
    if "--merge" in args:
        args.remove("--merge")
        if os.path.exists(out_path):
            with open(out_path) as f:
                kernels = json.load(f).get("kernels", {})
    for a in args:
        tag, path = a.split("=", 1)
        kernels[tag] = dict(source=os.path.basename(path), launches=read(path))
    doc = dict(kernels=kernels, how="ncu --set full --clock-control none, one steady-state step of tools/step_once.py "
                                    "(atari shape); cold-cache serialised launches: shares, not bench values")
    if dominant:
        tag, idx = dominant.split(":")
        l = kernels[tag]["launches"][int(idx)]
        doc["dominant_gemm_launch"] = dict(dram_bytes=l.get("dram_bytes"), duration_us=l.get("duration_us"),
                                           tensor_pipe_active_pct=l.get("tensor_pipe_active_pct"), grid=l["grid"],
                                           note=f"largest captured launch of {l['kernel']} (grid {l['grid']}) in "
                                                f"profiles/{kernels[tag]['source']}: dram read+write bytes per launch")
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)
    for tag, k in kernels.items():
        for l in k["launches"]:
            print(tag, {kk: (round(v, 2) if isinstance(v, float) else v) for kk, v in l.items()})


if __name__ == "__main__":
    main()
