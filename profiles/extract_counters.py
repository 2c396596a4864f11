#!/usr/bin/env python
"""ncu --csv metric log of ONE kernel launch -> per-update counters as JSON (read by bench.py).

    ncu --metrics <list below> -k regex:t6_replay -c 1 --csv --log-file gpurun_out/t6_counters.csv \
        python bench.py --tsteps 100 --steps 1 --warmup 3 --no-cpu --no-e2e --no-extra --no-config5
    python profiles/extract_counters.py gpurun_out/t6_counters.csv UPDATES profiles/r02_t6_counters.json

UPDATES = filter updates of the captured launch (filters x epochs)."""
import csv
import json
import sys

METRICS = ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,"
           "smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread")


def to_base(value, unit):
    v = float(str(value).replace(",", ""))
    u = (unit or "").lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
             "second": 1.0}
    return v * scale.get(u, 1.0)


def main():
    path, updates, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iname, iunit, ival, ikern = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("Kernel Name")
    m, kernel = {}, None
    for r in rows[1:]:
        m[r[iname]] = to_base(r[ival], r[iunit])
        kernel = r[ikern]
    dfma = m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
    dmul = m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
    dadd = m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
    dram = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    res = {"kernel": kernel, "updates_in_launch": updates, "source": f"profiles/{path.split('/')[-1]} (ncu, one launch)",
           "dfma_per_update": dfma / updates, "dmul_per_update": dmul / updates, "dadd_per_update": dadd / updates,
           "executed_flop_per_update": (2 * dfma + dmul + dadd) / updates,
           "dram_bytes_per_update": dram / updates, "dram_bytes_launch": dram,
           "warp_inst_per_update": m.get("smsp__inst_executed.sum", 0) / updates,
           "gpu_time_s_under_ncu": m.get("gpu__time_duration.sum"),
           "fp64_pipe_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "lanes_per_inst": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "registers": m.get("launch__registers_per_thread")}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
