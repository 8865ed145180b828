"""Development aid: run a script of this repository against a variant library (tools/dev/build_variant.py).
    python tools/dev/with_lib.py rolo_b200/lib/variants/<name>.so bench.py --steps 3 ..."""
import os, runpy, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
from rolo_b200 import build as B
lib = os.path.abspath(sys.argv[1])
B.build = lambda *a, **k: lib
B.LIB = lib
script = sys.argv[2]
sys.argv = sys.argv[2:]
runpy.run_path(script, run_name="__main__")
