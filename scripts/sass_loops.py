"""Static look at the loops of one kernel in libpdecg.so: for every backward branch, the opcode mix of the loop body and
the sum of the compiler's issue-stall counts (control field of each 128-bit SASS instruction) -- the minimum number of
issue cycles ONE warp needs per trip when it runs alone on its SM sub-partition (in-order issue).

    python scripts/sass_loops.py <mangled-kernel-name-substring> [min_body_instructions]
"""
import re
import subprocess
import sys

lib = "pdecontrolgym_b200/lib/libpdecg.so"
pat, min_body = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 100
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    lines = f.split("\n")
    ins, i = [], 0
    while i < len(lines):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            ins.append((int(m.group(1), 16), m.group(2), int(m2.group(1), 16) if m2 else 0))
            i += 2
        else:
            i += 1
    print(name, len(ins), "instructions")
    for addr, t, hi in ins:
        m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [(a, x, h) for a, x, h in ins if tgt <= a <= addr]
            if len(body) < min_body:
                continue
            ops, stalls = {}, 0
            for a, x, h in body:
                op = x.split()[1] if x.startswith("@") else x.split()[0]
                op = op.split(".")[0]
                ops[op] = ops.get(op, 0) + 1
                stalls += (h >> 41) & 0xF
            print(f"  loop {tgt:#x}..{addr:#x}: {len(body)} instr, stall-count sum {stalls}, "
                  + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda z: -z[1])[:8]))
