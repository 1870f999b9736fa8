"""Development: issue-slot cost of the backward's interior row body from the SASS, offline (no GPU).

    python tools/sass_cost.py [kernel-template-args, default "3,0,0,0,1,1"]

Compiles csrc/warp_bwd_tma.cu, extracts the kernel, finds the unrolled rows of the first interior_strip instantiation (by
their FADD2.RM floor instruction) and prints, per row, the instruction count and a slot-weighted cost with the weights
measured by tools/exp/exp_pipes.cu on B200 (FFMA / IADD = 1, IMAD / LOP3 = 1.4, packed f32x2 = 2.2, SHFL = 2.5, LDS / STS
0.6; drain blocks, which run once every few rows, are excluded)."""
import os, re, subprocess, sys, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1] if len(sys.argv) > 1 else "3,0,0,0,1,1"
key = "bwd_tma_kernelIfLi%sELb%sELb%sELb%sELb%sELb%sE" % tuple(args.split(","))   # the fp32 instantiation
obj = "/tmp/sass_cost_bwd.o"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-c",
                       os.environ.get("SASS_COST_SRC", os.path.join(root, "pwstablenet_b200/csrc/warp_bwd_tma.cu")), "-o", obj] + sys.argv[2:])
txt = subprocess.check_output(["cuobjdump", "-sass", obj], text=True)
lines, on = [], False
for l in txt.splitlines():
    if "Function :" in l:
        on = key in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            lines.append(m.group(2).strip())
W = {"IMAD": 1.4, "LOP3": 1.4, "FFMA2": 2.2, "FMUL2": 2.2, "FADD2": 2.2, "SHFL": 2.5, "LDS": 0.6, "STS": 0.6, "NOP": 0.0}
def op(l):
    t = l.split()
    if t[0].startswith("@"): t = t[1:]
    return t[0].split(".")[0]
starts = [i for i, l in enumerate(lines) if "FADD2.RM" in l or ("FADD.RM" in l and "12582912" in l)]
# group floor instructions into rows: scalar builds have two per row
rows = []
for i in starts:
    if not rows or i - rows[-1] > 40: rows.append(i)
print(f"kernel <{args}>: {len(lines)} instructions, {len(rows)} unrolled rows found")
for r in range(min(4, len(rows) - 1)):
    seg = lines[rows[r]:rows[r + 1]]
    # drop the drain blocks: from the LDS.128 of a pop to the last REDG that follows it
    keep, skip = [], 0
    for j, l in enumerate(seg):
        if "LDS.128" in l: skip = 1
        if not skip: keep.append(l)
        if skip and "REDG" in l and not any("REDG" in x for x in seg[j + 1:j + 4]): skip = 0
    c = collections.Counter(op(l) for l in keep)
    cost = sum(W.get(k, 1.0) * v for k, v in c.items())
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
    print(f"row {r}: {len(keep)} instr (of {len(seg)} static), weighted {cost:.0f} slots | {top}")
