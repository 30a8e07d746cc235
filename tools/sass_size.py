"""SASS size per kernel of a built library: python tools/sass_size.py lib.so [substring]"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
name, n, rows = None, 0, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name: rows.append((name, n))
        name, n = m.group(1), 0
    elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        n += 1
if name: rows.append((name, n))
for name, n in rows:
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if want in d:
        print(f"{n:7d} instr {n * 16 / 1024:7.1f} KB  {d[:90]}")
