#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report, one JSON record per launch (run where ncu is
installed; no GPU needed):  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic out.json]

--traffic writes the per-GEMM dram__bytes_read.sum + dram__bytes_write.sum record that bench.py
reports as roofline.traffic (profiles/r2_ncu_gemm_traffic.json), stamped with the current commit."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_pct",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]][:70]}
        for m, name in WANT.items():
            if m in idx:
                try:
                    v = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[m]]
                if name == "time_us":
                    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
                if name.endswith("_MB"):
                    v = v / 1e6 if u in ("byte",) else v * 1e-3 if u == "Kbyte" else v * 1e3 if u == "Gbyte" else v
                if name == "sm_ghz":
                    v = v / 1e9 if v > 1e6 else v / 1e3 if v > 100 else v
                rec[name] = round(v, 3)
        recs.append(rec)
    print(json.dumps(recs, indent=1))
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        gemms = [r for r in recs if "gemm" in r["kernel"]]
        per = [{"kernel": r["kernel"], "bytes": round((r.get("dram_read_MB", 0) + r.get("dram_write_MB", 0)) * 1e6)}
               for r in gemms]
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        json.dump({"source": f"ncu --set full --clock-control none, {rep}", "commit": commit,
                   "mean_bytes_per_launch": sum(p["bytes"] for p in per) / max(len(per), 1), "per_gemm": per},
                  open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
