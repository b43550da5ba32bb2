"""Dev aid: run a tools/ script against an alternative build of the library (timing experiments)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200"))
from mspde import _ffi
_ffi.LIB_PATH = os.path.abspath(sys.argv[1])
script = sys.argv[2]
sys.argv = sys.argv[2:]
exec(compile(open(script).read(), script, "exec"), {"__name__": "__main__", "__file__": os.path.abspath(script)})
