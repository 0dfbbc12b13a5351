#!/usr/bin/env python3
"""Static opcode histogram of one kernel in a cubin / .so (cuobjdump -sass). Usage: sass_mix.py <file> <regex> [--dump]"""
import collections, re, subprocess, sys

def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name: yield name, body
            name, body = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            body.append(line)
    if name: yield name, body

if __name__ == "__main__":
    path, rx = sys.argv[1], re.compile(sys.argv[2])
    for name, body in functions(path):
        if not rx.search(name): continue
        ops = collections.Counter()
        for l in body:
            t = re.sub(r"/\*[0-9a-fx]+\*/", "", l).strip().rstrip(";").split()
            if not t: continue
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += 1
        print(f"{name}: {len(body)} instructions")
        print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(45)))
        if "--dump" in sys.argv:
            for l in body: print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l))
