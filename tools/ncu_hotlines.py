"""Top CUDA source lines by warp-stall samples from an .ncu-rep captured with --import-source on (compile with -lineinfo).
Usage: python tools/ncu_hotlines.py rep.ncu-rep kernel_regex [topN]"""
import csv, io, subprocess, sys

def main(rep, kre, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None; cur_file = None; agg = {}
    for r in rows:
        if not r: continue
        if r[0] == "Kernel Name": print("##", r[1]); continue
        if r[0] == "File Name": cur_file = r[1].split("/")[-1]; continue
        if "# Samples" in r or "Warp Stall Sampling (All Samples)" in r:
            hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
        if hdr is None or len(r) < len(hdr): continue
        key_col = "Warp Stall Sampling (All Samples)" if "Warp Stall Sampling (All Samples)" in idx else "# Samples"
        try:
            s = float(r[idx[key_col]] or 0)
        except ValueError:
            continue
        if s <= 0: continue
        # cuda,sass view: rows carry a source line number + text in the first columns
        line = r[0]; text = r[1]
        k = (cur_file, line)
        a = agg.setdefault(k, [0.0, text, 0.0])
        a[0] += s
        if "stall_barrier" in idx:
            try: a[2] += float(r[idx["stall_barrier"]] or 0)
            except ValueError: pass
    tot = sum(v[0] for v in agg.values()) or 1
    print("total samples", tot)
    for (f, l), (s, text, bar) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*s/tot:5.1f}%  {f}:{l:>5s}  bar {100*bar/max(s,1):4.0f}% | {text.strip()[:120]}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
