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
        if m and int(m.group(1), 16) < int(addr, 16) and int(addr, 16) - int(m.group(1), 16) > 0x400:
            loops.append((int(m.group(1), 16), int(addr, 16)))
    # the main loop: the largest backward branch that contains no other one (the out-of-line mbarrier wait of the
    # pair-table kernels branches back over the whole function and is not a loop)
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    loop_start, loop_end = max(inner, key=lambda l: l[1] - l[0]) if inner else (None, None)
    if loop_start is None:
        print(name, "no loop found, total", len(lines)); continue
    body = [(int(a, 16), i) for a, i in lines if loop_start <= int(a, 16) <= loop_end]
    cold = []
    for a, ins in body:
        m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) - a > 0x800 and int(m.group(1), 16) <= loop_end:
            cold.append((a, int(m.group(1), 16)))
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
