"""Writes tests/golden/ref_dumps/: cluster files produced by the UNMODIFIED reference
(oracle/_ref/tri-L built from /root/reference/main.cpp) -- the byte-exact pins of the gap format
(main.cpp:232-257).  spin_* are deterministic functions of the lattice; fk_* use random bonds.
    python tests/golden/make_reference_dumps.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_dumps")
os.makedirs(OUT, exist_ok=True)
for L, n in ((4, 3), (8, 3), (16, 2), (32, 1)):
    base = tempfile.mkdtemp(prefix="bcref_")
    ref_runner.run_reference(L, 0.608, 1.966, 1.0, n, run=0, workdir=base)
    for i in range(n):
        shutil.copy(os.path.join(base, "r", "spin", str(L), "0", f"{i}.txt"), os.path.join(OUT, f"spin_{L}_{i}.txt"))
    shutil.copy(os.path.join(base, "r", "fk", str(L), "0", "0.txt"), os.path.join(OUT, f"fk_{L}_0.txt"))
    shutil.rmtree(base)
# the reference's own burn-file path: burn=1 writes {L}_burn.txt after generate_lattice + 1500*N steps
base = tempfile.mkdtemp(prefix="bcref_")
os.makedirs(os.path.join(base, "r", "burn"))
subprocess.run([ref_runner.binary(4), "0", "r", "1", "0", "0.608", "1.966", "1.0"], cwd=base, check=True,
               stdout=subprocess.DEVNULL)
shutil.copy(os.path.join(base, "r", "burn", "4_burn.txt"), os.path.join(OUT, "burn_4.txt"))
shutil.rmtree(base)
print(sorted(os.listdir(OUT)))
