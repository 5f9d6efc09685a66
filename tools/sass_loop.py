"""Summarise the hot loop of a kernel's SASS: instruction mix of the largest backward-branch loop.
usage: python tools/sass_loop.py <lib.so> <kernel-name-substring>"""
import collections
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((i - addr_index[tgt], addr_index[tgt], i))
    print(f"== {name[:150]}\n   {len(ins)} instructions, {len(loops)} backward branches")
    for n, s, e in sorted(loops, reverse=True)[:3]:
        cnt = collections.Counter()
        for _, t in ins[s:e + 1]:
            op = t.split()[0] if not t.startswith("@") else t.split()[1]
            cnt[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "LD", "ST", "SHFL", "BAR")) and "." in op else "")] += 1
        print(f"   loop {ins[s][0]:#x}..{ins[e][0]:#x}: {n + 1} instr: " + ", ".join(f"{k}:{v}" for k, v in cnt.most_common(24)))
