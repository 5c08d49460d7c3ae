"""Summarise an .ncu-rep capture (read here, no GPU needed) into the few numbers DESIGN.md / bench.py quote:
per kernel launch: duration, DRAM bytes read+written, DRAM throughput %, SM busy %, achieved occupancy, registers, and the top
stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof_be.ncu-rep > profiles/r01_prof_be.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__cycles_elapsed.avg.per_second"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("no data in", path)
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in hdr if "warp_issue_stalled" in h and h.endswith(".pct")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"== {name}  (id {r[idx['ID']]})")
        for w in WANT:
            if w in idx:
                print(f"   {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            u = units[idx["dram__bytes_read.sum"]]
            print(f"   {'traffic = dram read + write':70s} {rd + wr:16.3f} {u}")
        except Exception:
            pass
        st = []
        for c in stall_cols:
            try:
                st.append((float(r[idx[c]].replace(",", "")), c))
            except Exception:
                pass
        for v, c in sorted(st, reverse=True)[:5]:
            print(f"   stall {c:64s} {v:16.3f}")
    print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
