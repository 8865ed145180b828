"""Development aid: build a variant of the library with extra nvcc flags for ONE source file (tuning constants), next to the product
library: rolo_b200/lib/variants/<name>.so.   python tools/dev/build_variant.py <name> <source.cu> -DROLO_KNN_RAW_CAP=2048 ..."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rolo_b200 import build as B

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
out_dir = os.path.join(B.OUT_DIR, "variants"); os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"{name}_{os.path.splitext(src)[0]}.o")
subprocess.check_call([B.NVCC, *B.ARCH, *[f for f in B.COMMON if f not in ("-Xptxas", "-v")], *B.PER_FILE.get(src, []), *flags, "-c", os.path.join(B.SRC, src), "-o", obj])
objs = [obj if s == src else os.path.join(B.OBJ_DIR, os.path.splitext(s)[0] + ".o") for s in B.SOURCES]
lib = os.path.join(out_dir, f"{name}.so")
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-ccbin", B.HOST_CXX, "-o", lib, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
print(lib)
