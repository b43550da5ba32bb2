import sys, os, ctypes
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200"))
import torch
from mspde import _ffi
L = _ffi.lib(); dev = "cuda"
m = n = 8192; k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A = torch.randn(k, m, dtype=torch.complex128, device=dev); B = torch.randn(k, n, dtype=torch.complex128, device=dev)
C = torch.zeros(m, n, dtype=torch.complex128, device=dev)
def run():
    L.mspde_zgemm_tn(_ffi.current_stream(), 1, m, n, k, ctypes.c_double(-1.0), _ffi.ptr(A), A.stride(0), _ffi.ptr(B), B.stride(0),
                     ctypes.c_double(1.0), _ffi.ptr(C), C.stride(0), 0, None, 1)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
