"""Probe kept from round 2: compares libvdd's in-kernel Exp(1) stream with torch.empty(n).exponential_(1) (0 mismatches expected).
The same probe with a switchable float recipe showed that ATen's CUDA exponential uses the fast __logf: with logf 85 % of the
elements differ in the last bits, with __logf none."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.getcwd())
from llava_align_b200 import _lib as L
lib = L.load()
torch.cuda.init(); torch.zeros(1, device="cuda"); gen = torch.cuda.default_generators[0]
for var in range(1):
    torch.manual_seed(5); torch.rand(3, device="cuda")
    seed, off = gen.initial_seed(), gen.get_offset()
    ref = torch.empty(4, 32000, device="cuda").exponential_(1)
    got = torch.empty(4, 32000, device="cuda")
    L.check(lib.vdd_debug_philox_exponential(C.c_uint64(seed), C.c_uint64(off), got.numel(), C.c_void_p(got.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "p")
    torch.cuda.synchronize()
    print("variant", var, "mismatches", int((got != ref).sum()), "max rel", float(((got-ref).abs()/ref).max()))
