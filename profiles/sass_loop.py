"""Instruction count and opcode histogram of the innermost loop (the smallest backward-branch span that contains MUFU.EX2) of a kernel
in the built library: python profiles/sass_loop.py <mangled-name-substring> [library].  Static view; the dynamic count is ncu's."""
import collections
import re
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "litegs_b200/liblitegs_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    loops = []
    for addr, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?|BRA(?:\.\w+)*\s+0x([0-9a-f]+)", t)
        m2 = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt <= addr:
                body = [x for x in ins if tgt <= x[0] <= addr]
                if any("MUFU.EX2" in b[1] for b in body):
                    loops.append(body)
    if not loops:
        print(name, "no loop found", len(ins)); continue
    body = min(loops, key=len)
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", b[1]).split()[0].split(".")[0] for b in body)
    print(f"{name}\n  total {len(ins)}  innermost EX2 loop {len(body)} instructions: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
