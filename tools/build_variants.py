"""Builds experimental variants of libnz_engine.so next to the default (gpurun_out is not shipped: they go to build_variants/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
from nuzero_b200 import build as b
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "build_variants")
os.makedirs(out, exist_ok=True)
variants = {}
for arg in sys.argv[1:]:
    name, _, defs = arg.partition("=")
    variants[name] = [d for d in defs.split(",") if d]
def one(item):
    name, defs = item
    path = os.path.join(out, "libnz_%s.so" % name)
    b.build(defines=defs, out=path)
    return path
with ThreadPoolExecutor(8) as ex:
    for p in ex.map(one, variants.items()):
        print(p)
