"""Turn gpurun_out/*.ncu-rep into committed summaries under profiles/ (run on the CPU box: `ncu -i` needs no GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    header, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        recs.append({h: (v, u) for h, v, u in zip(header, r, units)})
    return recs


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    dst = sys.argv[2] if len(sys.argv) > 2 else "profiles"
    os.makedirs(dst, exist_ok=True)
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".ncu-rep"):
            continue
        recs = raw_rows(os.path.join(src, fn))
        lines = ["# %s" % fn, "", "`ncu --set full --clock-control none --import-source on` on one B200 (see bench/ncu_capture.sh).",
                 "Roofline denominators (MEASURED_PEAKS.json): HBM %s GB/s, bf16 %s TFLOP/s (burst)." % (
                     peaks.get("hbm_gbs"), peaks.get("bf16_tflops")), ""]
        for rec in recs:
            name = rec.get("Kernel Name", ("?", ""))[0]
            lines.append("## `%s`" % name[:140])
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for k in KEYS:
                if k in rec:
                    lines.append("| %s | %s | %s |" % (k, rec[k][0], rec[k][1]))
            try:
                BY = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                TM = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0,
                      "second": 1.0}

                def qty(key, table):           # every metric carries its OWN unit (read and write often differ)
                    val, unit = rec[key][0], rec[key][1]
                    if unit not in table:
                        raise KeyError(unit)
                    return float(val.replace(",", "")) * table[unit]

                secs = qty("gpu__time_duration.sum", TM)
                nbytes = qty("dram__bytes_read.sum", BY) + qty("dram__bytes_write.sum", BY)
                gbs = nbytes / secs / 1e9
                peak = float(peaks.get("hbm_gbs", 6478.3))
                if 0.0 < gbs < 20.0 * peak:
                    lines.append("| derived DRAM GB/s | %.1f (%.3f of measured copy peak %.0f GB/s) | |" % (gbs, gbs / peak, peak))
            except Exception:
                pass
            lines.append("")
        with open(os.path.join(dst, fn.replace(".ncu-rep", ".md")), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote", fn, len(recs), "kernels")


if __name__ == "__main__":
    main()
