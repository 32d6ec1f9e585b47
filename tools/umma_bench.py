"""Microbenchmark: cycles per tcgen05.mma (M=128, N, K=16 bf16) as issued by this library's helpers."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from context_attentive_ir_b200 import lib
L = C.CDLL(lib.LIB_PATH)
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for uniform in (0, 1):
    for N in (16, 32, 64, 96, 128, 256):
        for K in (64,):
            reps = 200
            L.cair_umma_bench(N, K, reps, uniform, C.c_void_p(out.data_ptr()), None)
            torch.cuda.synchronize()
            n = reps * K // 16
            print('uniform=%d N=%3d: issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA (floor %.0f)' % (
                uniform, N, out[0].item() / n, out[1].item() / n, 128 * N / 256))
