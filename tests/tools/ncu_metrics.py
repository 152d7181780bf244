"""Build-container tool: turn an `ncu --set full` capture (gpurun_out/*.ncu-rep) into the small JSON that
profiles/ keeps and bench.py reads for `roofline.traffic`.
    python -m tests.tools.ncu_metrics gpurun_out/<capture>.ncu-rep profiles/<name>.json [steps_per_launch]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size"]

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
        "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m in WANT:
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[m] = v * UNIT.get(units[i], 1.0) if units[i] in UNIT else v
        launches.append(d)
    res = {"source": rep, "steps_per_launch": steps, "launches": launches}
    if launches:
        l0 = launches[-1]
        rd, wr = l0.get("dram__bytes_read.sum"), l0.get("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            res["dram_bytes_per_step"] = (rd + wr) / steps
        if "gpu__time_duration.sum" in l0:
            res["us_per_step_under_ncu"] = l0["gpu__time_duration.sum"] / steps * 1e6
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "launches"}, indent=1))


if __name__ == "__main__":
    main()
