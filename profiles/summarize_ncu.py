#!/usr/bin/env python
"""Turn an `ncu --set full` report into the text summary kept under profiles/ and (with --traffic) into
profiles/ncu_traffic.json: DRAM bytes per bench-table kernel name for ONE step, each stamped with the digest of the CUDA
sources that kernel is compiled from (bench.py only attaches a figure while its digest still matches; kernels absent from a
partial re-capture keep their entry as long as their own sources did not move).

    python profiles/summarize_ncu.py gpurun_out/r02_ncu_full.ncu-rep profiles/r02_ncu_full_summary.txt [--traffic profiles/ncu_traffic.json]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WANT = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
NAMES = [("k_gather", "final_gather"), ("k_final", "final_gather"), ("k_psnr_ssim", "psnr_ssim"), ("k_down_u8", "down_u8"),
         ("k_down_u16", "down_f32"), ("k_down_f32s", "down_f32"), ("k_up_level", "up_level"), ("k_iqa_reduce", "iqa_reduce")]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1.0, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6}


def main():
    rep, out_txt = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    stalls = [h for h in head if "issue_stalled" in h and "per_issue_active" in h]
    lines, traffic = [], {}
    steps_seen = 0  # a step starts with k_down_u8: the traffic table sums the FIRST step only (the capture may run into the next)
    for r in body:
        name = r[col["Kernel Name"]]
        if "k_down_u8" in name:
            steps_seen += 1
        lines.append(f"Kernel Name = {name}")
        for w in WANT + stalls:
            if w in col:
                lines.append(f"{w} = {r[col[w]]} {units[col[w]]}")
        lines.append("-----")
        table = next((t for k, t in NAMES if k in name), None)
        if "k_down_f32s<1>" in name.replace("(bool)", "").replace(" ", ""):
            table = None  # weight-pyramid launches: first call only (cached afterwards), not part of a steady-state step
        if table and steps_seen <= 1:
            def val(metric):
                return float(r[col[metric]].replace(",", "")) * UNIT.get(units[col[metric]], 1.0)
            t = traffic.setdefault(table, {"dram_bytes_per_step": 0.0, "ncu_time_ms": 0.0, "launches": 0})
            t["dram_bytes_per_step"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            t["ncu_time_ms"] += val("gpu__time_duration.sum")
            t["launches"] += 1
    with open(out_txt, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{out_txt}: {len(body)} launches")
    if traffic_path:
        import bench
        # every kernel carries the digest of ITS sources; entries of an earlier capture are kept while theirs still matches
        # (a partial re-capture after a change to one translation unit), and say which summary they come from
        doc = {"kernels": {}}
        if os.path.exists(traffic_path):
            with open(traffic_path) as f:
                doc = json.load(f)
        kept = {k: v for k, v in doc.get("kernels", {}).items() if k not in traffic and v.get("src_digest") == bench.csrc_digest(k)}
        for k, v in traffic.items():
            v["src_digest"] = bench.csrc_digest(k)
            v["summary"] = os.path.basename(out_txt)
        doc = {"source": "dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of ONE step that the bench table files under each "
                         "name (ncu --set full --clock-control none, cfg4_200mp, 1 GPU); per kernel: the summary file it was read from and the "
                         "digest of the sources it is compiled from",
               "workload": "cfg4_200mp", "kernels": {**kept, **traffic}}
        with open(traffic_path, "w") as f:
            json.dump(doc, f, indent=1)
        print(f"{traffic_path}: captured {list(traffic)}, kept {list(kept)}")

if __name__ == "__main__":
    main()
