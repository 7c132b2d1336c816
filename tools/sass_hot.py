"""Static look at k_fused_fast's unchecked 4-record block: instruction and MOV counts (no GPU needed).
usage: sass_hot.py [libogpu.so] [instance substring]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "opengemini_b200/libogpu.so"
inst = sys.argv[2] if len(sys.argv) > 2 else "k_fused_fastILi11ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
body = next(b for b in blocks if inst in b.split("\n")[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = next(i for i, (a, t) in enumerate(ins) if "CREDUX" in t)
tgt = None
for a, t in ins[idx:idx + 12]:
    m = re.search(r"@P\d\s+BRA\s+0x([0-9a-f]+)", t)
    if m: tgt = int(m.group(1), 16); break
start = next(i for i, (a, t) in enumerate(ins) if a == tgt)
end = next(i for i in range(start, len(ins)) if ins[i][1].startswith("VOTE.ALL") or "BRA.U" in ins[i][1] and i > start + 50)
blk = ins[start:end]
# hot = instructions not inside slow paths: approximate by excluding ranges between a taken-forward "@!P BRA" and its target? keep simple: count all + key opcodes
ops = {}
for a, t in blk:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
print(f"unchecked block: {len(blk)} instrs (4 records, slow paths included), regs see ptxas -v")
print(sorted(ops.items(), key=lambda x: -x[1])[:14])
mov = sum(1 for a, t in blk if re.search(r"(IMAD\.MOV|\bMOV\b)", t))
print("moves:", mov)
