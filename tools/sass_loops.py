"""Static SASS view of a kernel's loops: for every backward branch of the chosen kernel in an object file, the span's
instruction count and opcode histogram (the tile loop of the update kernels is the span holding the UTCHMMA issue).

    python tools/sass_loops.py <file.o|.so> <kernel name substring> [min span]
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def op(src):
    t = src.split()
    o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]


if __name__ == "__main__":
    path, sub = sys.argv[1], sys.argv[2]
    min_span = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    for name, ins in kernels(path).items():
        if sub not in name:
            continue
        print(name, len(ins), "instructions")
        addr = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, s) in enumerate(ins):
            m = re.search(r"BRA.*?(0x[0-9a-f]+)", s)
            if not m:
                continue
            t = int(m.group(1), 16)
            if t < a and t in addr and i - addr[t] >= min_span:
                span = ins[addr[t]:i + 1]
                c = collections.Counter(op(x) for _, x in span)
                print(f"  loop {t:#x}..{a:#x}: {len(span)} instructions")
                print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common(28)))
