"""Static SASS statistics of one kernel of libbfa_b200.so: opcode histogram of its longest branch-free stretches
(the unrolled 8-frame body of the banded Viterbi kernel is the longest one)."""
import re, subprocess, sys, collections
lib = "bournemouth-forced-aligner_b200/lib/libbfa_b200.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "viterbi_band3_kernelILi66ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
body = [f for f in funcs if f.startswith("_Z") and pat in f.split("\n")[0]]
assert body, "kernel not found"
lines = body[0].split("\n")
print("kernel", lines[0])
ins = []
for l in lines:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        txt = m.group(2).strip()
        ops = txt.split()
        op = ops[1] if ops[0].startswith("@") else ops[0]
        ins.append((int(m.group(1), 16), op, txt))
print("total instructions", len(ins))
# split into branch-free stretches
stretches, cur = [], []
for a, op, txt in ins:
    cur.append((a, op, txt))
    if op.split(".")[0] in ("BRA", "EXIT", "RET", "BSYNC", "BSSY", "WARPSYNC", "CALL"):
        stretches.append(cur); cur = []
stretches.append(cur)
stretches.sort(key=len, reverse=True)
for s in stretches[:int(sys.argv[2]) if len(sys.argv) > 2 else 3]:
    h = collections.Counter(op.split(".")[0] for _, op, _ in s)
    print(f"\nstretch {s[0][0]:#x}..{s[-1][0]:#x}: {len(s)} instructions")
    print("  " + ", ".join(f"{k}:{v}" for k, v in h.most_common()))
