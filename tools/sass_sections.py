#!/usr/bin/env python
"""SASS size of one kernel by source section:  python tools/sass_sections.py <object.o> <mangled-name-substring> <source.cu> marker1 marker2 ...
(markers are substrings of source lines; a SASS instruction belongs to the last marker at or before its line)."""
import re, subprocess, sys, tempfile, os, glob

def main():
    obj, sub, srcf = sys.argv[1:4]
    marks_txt = sys.argv[4:]
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = glob.glob(os.path.join(d, "*.cubin"))[0]
    lines = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
    start = end = None
    for i, l in enumerate(lines):
        if l.startswith("//--------------------- .text.") and sub in l and start is None:
            start = i
        elif start is not None and l.startswith("//--------------------- ") and i > start:
            end = i
            break
    src = open(srcf).read().split("\n")
    marks = []
    for t in marks_txt:
        for k, l in enumerate(src):
            if t in l:
                marks.append((t, k + 1))
                break
    cur, counts, n = None, {}, 0
    for ln in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln) and cur:
            counts[cur] = counts.get(cur, 0) + 1
            n += 1
    sec = {}
    for (f, l), c in counts.items():
        if f != os.path.basename(srcf):
            s = "inlined: " + f
        else:
            s = "(before first marker)"
            for name, m in marks:
                if l >= m:
                    s = name
        sec[s] = sec.get(s, 0) + c
    for s, c in sorted(sec.items(), key=lambda kv: -kv[1]):
        print("%-60s %6d instr %6.1f KB" % (s[:60], c, c * 16 / 1024))
    print("total %d instr %.1f KB" % (n, n * 16 / 1024))

if __name__ == "__main__":
    main()
