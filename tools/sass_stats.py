"""Instruction histogram of the main loop of a kernel in libbc_engine.so (development aid)."""
import re, collections, subprocess, sys
import os
lib = os.environ.get("BC_ENGINE_LIB", "blume_capel_b200/libbc_engine.so")
pat = sys.argv[1] if len(sys.argv) > 1 else "fast_kernelILi0ELb0ELb0ELi1"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ALU = {"LOP3","PRMT","SHF","ISETP","SEL","IADD3","VIADD","LEA","IABS","VIMNMX","MOV","PLOP3","P2R","R2P","BMSK","SGXT","FLO","POPC","IADD","UIADD3"}
FMA = {"IMAD","IDP","FFMA","FMUL","FADD","HFMA2"}
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if pat not in name: continue
    lines = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
    loops = []
    for addr, ins in lines:
        m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) < int(addr, 16) and int(addr, 16) - int(m.group(1), 16) > 0x40:
            loops.append((int(m.group(1), 16), int(addr, 16)))
    # the main loop: the backward branch that encloses the most 16-byte row stores (the rolled loops of the cold tie
    # path are loops too, and the out-of-line mbarrier wait of the pair-table kernels branches back over everything)
    def n_stores(l):
        return sum(1 for a, i in lines if l[0] <= int(a, 16) <= l[1] and "STG.E.128" in i)
    cand = [l for l in loops if n_stores(l) > 0]
    cand = [l for l in cand if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in cand)]  # innermost of those
    loop_start, loop_end = max(cand, key=n_stores) if cand else (None, None)
    if loop_start is None:
        print(name, "no loop found, total", len(lines)); continue
    body = [(int(a, 16), i) for a, i in lines if loop_start <= int(a, 16) <= loop_end]
    # cold regions: forward branches over a region that holds a loop of its own (the tie repair)
    cold = []
    for a, ins in body:
        m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) > a and int(m.group(1), 16) <= loop_end:
            t = int(m.group(1), 16)
            if any(a < l[0] and l[1] < t for l in loops) or t - a > 0x800:
                cold.append((a, t))
    cold = [c for c in cold if not any(o != c and o[0] <= c[0] and c[1] <= o[1] for o in cold)]
    body = [(a, i) for a, i in body if not any(c0 < a < c1 for c0, c1 in cold)]
    print(" cold regions skipped:", [(hex(a), hex(b)) for a, b in cold])
    ops = collections.Counter(); full = collections.Counter()
    for a, ins in body:
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        full[ins.split()[0]] += 1
        ops[ins.split()[0].split(".")[0]] += 1
    n = len(body)
    rows = max(1, sum(v for k, v in full.items() if k.startswith("STG.E.128")))  # one 16-byte store = one row of 16 sites
    sites = 16 * rows
    alu = sum(v for k, v in ops.items() if k in ALU); fma = sum(v for k, v in ops.items() if k in FMA)
    print(name); print(" main loop %#x..%#x: %d instructions per %d rows = %d sites -> %.1f instr/site; ALU-pipe %d (%.2f/site) FMA-pipe %d (%.2f/site) LDS %d LDGSTS %d"
          % (loop_start, loop_end, n, rows, sites, n / sites, alu, alu / sites, fma, fma / sites, ops["LDS"], ops["LDGSTS"]))
    print(" ", sorted(full.items(), key=lambda x: -x[1])[:24])
