"""Per-kernel SASS fingerprint of libbc_engine.so: registers, shared memory, instruction count and an md5 of the
instruction text (addresses and encodings stripped).  Used to prove that an edit left a tuned kernel untouched
(the streaming kernel sits at a register cliff) and to commit opcode evidence under profiles/.
    python tools/sass_fingerprint.py [lib] [substring]"""
import collections
import hashlib
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "blume_capel_b200", "libbc_engine.so")
pat = sys.argv[2] if len(sys.argv) > 2 else ""
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = []
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", f)
    names.append((name, ins))
dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in names), capture_output=True, text=True).stdout.split("\n")
for (name, ins), d in sorted(zip(names, dem), key=lambda x: x[1]):
    if pat and pat not in d and pat not in name:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
    h = hashlib.md5("\n".join(re.sub(r"0x[0-9a-f]+", "X", i) for i in ins).encode()).hexdigest()[:12]
    reg, stack, sh, loc = usage.get(name, (-1, -1, -1, -1))
    d = d.replace("bc::", "").replace("(SweepParams)", "").replace("void ", "")
    print(f"{d:70s} reg={reg:3d} stack={stack:3d} smem={sh:6d} n={len(ins):5d} md5={h} "
          f"IMAD={ops['IMAD']} LOP3={ops['LOP3']} PRMT={ops['PRMT']} IDP={ops['IDP']} LDS={ops['LDS']} STL={ops['STL']} LDL={ops['LDL']}")
