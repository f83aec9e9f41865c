import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from phendiff_b200 import _lib
L = _lib.lib()
n, s, c = 64, 1024, 512
qkv = (torch.randn(n, s, 3 * c) * 0.5).half().cuda()
out = torch.empty(n, s, c, dtype=torch.half, device="cuda")
for _ in range(2):
    _lib.check(L.pd_test_attention(7, 2, n, s, c, 8, C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), None))
torch.cuda.synchronize()
