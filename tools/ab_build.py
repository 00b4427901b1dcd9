"""Build tuning variants of the library side by side (same ABI) for same-box A/B timing:
   python tools/ab_build.py tag1:-DFOO tag2:-DBAR,-DBAZ  ->  coral_b200/lib/ab/libcoral_b200_<tag>.so
Run one with CORAL_B200_LIB=coral_b200/lib/ab/libcoral_b200_<tag>.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coral_b200 import _build as b
out_dir = os.path.join(b.LIB_DIR, "ab")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    tag, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj_dir = os.path.join(b.OBJ_DIR, "ab_" + tag)
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src in b.CU_SOURCES + b.CC_SOURCES:
        op = os.path.join(obj_dir, src + ".o")
        objs.append(op)
        cmd = [b._nvcc(), *b.NVCC_FLAGS, *flags, "-x", "cu", "-c", os.path.join(b.CSRC, src), "-o", op]
        procs.append(subprocess.Popen(cmd))
    assert all(p.wait() == 0 for p in procs)
    lib = os.path.join(out_dir, f"libcoral_b200_{tag}.so")
    subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs])
    print(lib)
