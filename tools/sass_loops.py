#!/usr/bin/env python
"""Static look at a kernel's SASS: every backward branch closes a loop; print each loop's address range and its
instruction mix by opcode family.  Usage: cuobjdump -sass x.o | python tools/sass_loops.py <kernel-name-substring>"""
import collections
import re
import sys

want = sys.argv[1] if len(sys.argv) > 1 else ""
ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
cur, funcs = None, {}
for line in sys.stdin:
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
        funcs[cur] = []
        continue
    m = ins_re.match(line)
    if m and cur is not None:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))


def family(txt):
    t = txt.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("IMAD", "LDS", "STS", "LDG", "STG", "BAR", "MUFU")) and "." in op else "")


for name, ins in funcs.items():
    if want not in name:
        continue
    print("==", name, len(ins), "instructions")
    total = collections.Counter(family(t) for _, t in ins)
    loops = []
    for addr, txt in ins:
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                loops.append((tgt, addr))
    for lo, hi in sorted(loops, key=lambda x: (x[0], -x[1])):
        body = [t for a, t in ins if lo <= a <= hi]
        c = collections.Counter(family(t) for t in body)
        print("  loop 0x%05x..0x%05x  %5d instr : %s" % (lo, hi, len(body), ", ".join("%s %d" % kv for kv in c.most_common(28))))
