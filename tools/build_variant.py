"""Build a tuning variant of the library: tools/build_variant.py NAME -DNEDDF_TC_LOAD_WARPS=12 ...
-> neddf_b200/variants/libneddf_b200_NAME.so; run anything with NEDDF_B200_LIB=<that path>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

name, defs = sys.argv[1], sys.argv[2:]
g.build()
out_dir = os.path.join(ROOT, "neddf_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"field_tc_{name}.o")
subprocess.check_call([g._nvcc()] + g.NVCC_FLAGS + defs + ["-Xptxas", "-v", "-c", os.path.join(g.CSRC, "field_tc.cu"), "-o", obj])
objs = [os.path.join(g.CSRC, s.replace(".cu", ".o")) for s in g.SOURCES if s != "field_tc.cu"] + [obj]
lib = os.path.join(out_dir, f"libneddf_b200_{name}.so")
subprocess.check_call([g._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
print(lib)
