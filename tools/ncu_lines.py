#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]
(needs kernels compiled with -lineinfo and a capture with --import-source on)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fn, fpath, hdr = None, None, None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            fn = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] not in ("", "-"):
            if want and want not in fn:
                continue
            ci = {n: i for i, n in enumerate(hdr)}
            def val(name):
                try:
                    return float(r[ci[name]])
                except (ValueError, KeyError):
                    return 0.0
            d = agg.setdefault(fn, {}).setdefault((fpath, r[0], r[1].strip()), {"samples": 0.0, "inst": 0.0, "stalls": {}})
            d["samples"] += val("Warp Stall Sampling (All Samples)")
            d["inst"] += val("Instructions Executed")
            for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait", "stall_lg", "stall_membar",
                      "stall_math", "stall_not_selected", "stall_branch_resolving", "stall_no_inst", "stall_sleep"):
                d["stalls"][k] = d["stalls"].get(k, 0.0) + val(k)
    for fn, lines in agg.items():
        tot = sum(d["samples"] for d in lines.values()) or 1.0
        toti = sum(d["inst"] for d in lines.values()) or 1.0
        print("===== %s   (samples %d, warp-instructions %d)" % (fn[:90], tot, toti))
        st = {}
        for d in lines.values():
            for k, v in d["stalls"].items():
                st[k] = st.get(k, 0.0) + v
        print("   stall mix: " + ", ".join("%s %.0f%%" % (k[6:], 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.02))
        for (fp, ln, src), d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
            big = ", ".join("%s %.0f" % (k[6:], v) for k, v in sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:2] if v > 0)
            print("%5.1f%% smp %5.1f%% inst  %s:%s  %-80s [%s]" % (100 * d["samples"] / tot, 100 * d["inst"] / toti, fp, ln, src[:80], big))


if __name__ == "__main__":
    main()
