"""Builds experiment variants of one kernel file: the library relinked with that file compiled under extra -D flags.
    python profiles/variants.py field_bin.cu name1:-DSD_TB_NRA=3,-DSD_TB_STAGE=4096 name2:...
writes scenedino_b200/build/variants/lib_<name>.so; run with SD_B200_LIB=<that path>."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scenedino_b200 import build as B  # noqa: E402

B.build_library()
src = os.path.join(B.CSRC, sys.argv[1])
vdir = os.path.join(B.HERE, "build", "variants")
os.makedirs(vdir, exist_ok=True)
objs = [os.path.join(B.HERE, "build", os.path.basename(s)[:-3] + ".o") for s in B.sources() if s != src]
procs = []
for spec in sys.argv[2:]:
    name, _, flags = spec.partition(":")
    obj = os.path.join(vdir, f"{os.path.basename(src)[:-3]}_{name}.o")
    cmd = [B._nvcc(), *B.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-c", src, "-o", obj]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, p in procs:
    out, _ = p.communicate()
    if p.returncode:
        print(f"--- {name} FAILED\n{out}")
        continue
    if out.strip():
        print(f"--- {name}\n{out}")
    lib = os.path.join(vdir, f"lib_{name}.so")
    subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj, *objs], check=True)
    print("built", lib)
