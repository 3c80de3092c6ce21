"""Per-CUDA-source-line hot spots of an .ncu-rep (needs -lineinfo and --import-source on).
usage: python profiles/source_hotspots.py <rep> [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    fname = ""
    agg = {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            g = lambda name: float(r[hdr.index(name)] or 0) if name in hdr else 0.0
            key = (fname, int(r[0]))
            v = agg.setdefault(key, dict(src=r[1].strip(), samples=0, inst=0, long_sb=0, short_sb=0, barrier=0, lg=0,
                                         mio=0, wait=0, math=0))
            v["samples"] += g("# Samples")
            v["inst"] += g("Instructions Executed")
            for k, n in (("long_sb", "stall_long_sb"), ("short_sb", "stall_short_sb"), ("barrier", "stall_barrier"),
                         ("lg", "stall_lg"), ("mio", "stall_mio"), ("wait", "stall_wait"), ("math", "stall_math")):
                v[k] += g(n)
    ts = sum(v["samples"] for v in agg.values()) or 1
    ti = sum(v["inst"] for v in agg.values()) or 1
    print(f"# {rep}: total samples {ts:.0f}, warp instructions {ti:.0f}")
    print("# samp%  inst%  long_sb short_sb barrier lg mio wait math | file:line source")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print(f"{v['samples'] / ts * 100:5.1f} {v['inst'] / ti * 100:6.1f}  {v['long_sb']:6.0f} {v['short_sb']:6.0f} "
              f"{v['barrier']:6.0f} {v['lg']:4.0f} {v['mio']:4.0f} {v['wait']:5.0f} {v['math']:4.0f} | {f}:{ln} {v['src'][:110]}")


if __name__ == "__main__":
    main()
