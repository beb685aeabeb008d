#!/usr/bin/env python3
"""Dynamic instruction count per Poseidon2 permutation from the SASS of the microbenchmark's k_perm<256,3>: the kernel is straight-line
code plus counted loops (backward branches); each backward branch closes a loop whose trip count is given on the command line in
order of appearance of the loop END (4, 21, 4 for the default build: first full rounds, internal rounds, last full rounds; the
outer `iters` loop is the last backward branch and gets weight 1).

    python tools/sass_dyn.py build/var/mb_x [trip counts ...]
Prints per-opcode dynamic counts, the multiplier-pipe cycles (IMAD 2, IMAD.HI / IMAD.WIDE 4) and ALU-pipe instruction count."""
import re, subprocess, sys
from collections import Counter

exe = sys.argv[1]
trips = [int(x) for x in sys.argv[2:]] or [4, 21, 4]
kern = "k_permILi256ELi3"
out = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
body = [b for b in blocks if kern in b.split("\n")[0]][0]
ins = []
for ln in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
weight = [1] * len(ins)
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
    if "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
# innermost loops get their trip counts in order of loop end; the outermost (iters) loop weight 1
loops.sort(key=lambda x: x[1])
inner = [l for l in loops if not any(o[0] <= l[0] and l[1] < o[1] and o != l for o in loops if False)]
for k, (lo, hi) in enumerate(loops):
    w = trips[k] if k < len(trips) else 1
    for j in range(lo, hi + 1):
        weight[j] *= w
# only count inside the outermost loop (the iters loop) if present
if loops:
    olo, ohi = min(l[0] for l in loops), max(l[1] for l in loops)
    outer = [l for l in loops if l[0] == olo and l[1] == ohi]
    rng = range(olo, ohi + 1) if outer else range(len(ins))
else:
    rng = range(len(ins))
cnt = Counter()
for j in rng:
    op = ins[j][1].split()[0]
    if op.startswith("@"):
        op = ins[j][1].split()[1]
    cnt[op] += weight[j]
tot = sum(cnt.values())
fma = sum(v * (4 if ("IMAD.HI" in k or "IMAD.WIDE" in k) else 2) for k, v in cnt.items() if k.startswith("IMAD"))
alu = sum(v for k, v in cnt.items() if not k.startswith("IMAD"))
wide = sum(v for k, v in cnt.items() if "IMAD.HI" in k or "IMAD.WIDE" in k)
print("%s loops=%s total=%d fma_cycles=%d alu_instr=%d wide+hi=%d units=%d" % (exe, [(b - a + 1) for a, b in loops], tot, fma, alu, wide, fma + 2 * alu + 2 * wide))
print("  " + "  ".join("%s:%d" % kv for kv in cnt.most_common(12)))
