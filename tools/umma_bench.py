"""Microbenchmark: cycles per tcgen05.mma (M=128, N, K=16 bf16) as issued by this library's helpers."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from context_attentive_ir_b200 import lib
L = C.CDLL(lib.LIB_PATH)
out = torch.zeros(8, dtype=torch.int64, device='cuda')
for nw in (1, 2, 4):
    for N in (16, 32, 64, 96, 128):
        if nw * N > 512:
            continue
        reps, K = 200, 64
        L.cair_umma_bench(N, K, reps, 1 | (nw << 4), C.c_void_p(out.data_ptr()), None)
        torch.cuda.synchronize()
        n = reps * K // 16
        o = out.cpu().tolist()
        print('warps=%d N=%3d: per-warp issue %.1f cyc/MMA, all retired after %.1f cyc per (MMA of one warp) => %.1f cyc/MMA aggregate (floor %.0f)' % (
            nw, N, o[0] / n, max(o[1:2 * nw:2]) / n, max(o[1:2 * nw:2]) / n / nw, 128 * N / 256))

print('row-shifted A descriptor (start address + 16*shift bytes), N=96, one warp:')
for shift in range(8):
    reps, K = 200, 64
    L.cair_umma_bench(96, K, reps, 1 | (1 << 4) | (shift << 8), C.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    n = reps * K // 16
    o = out.cpu().tolist()
    print('  shift=%d: issue %.1f cyc/MMA, retired after %.1f cyc/MMA' % (shift, o[0] / n, o[1] / n))

print('tcgen05.commit every P MMAs (N=96, one warp):')
for P in (0, 36, 18, 9, 3, 1):
    reps, K = 200, 64
    L.cair_umma_bench(96, K, reps, 1 | (1 << 4) | (P << 12), C.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    n = reps * K // 16
    o = out.cpu().tolist()
    print('  P=%2d: issue %.1f cyc/MMA, retired after %.1f cyc/MMA' % (P, o[0] / n, o[1] / n))

print('A operand from tensor memory (.ts form):')
for N in (16, 32, 64, 96, 128, 256):
    reps, K = 200, 64
    L.cair_umma_bench(N, K, reps, 1 | 2 | (1 << 4), C.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    n = reps * K // 16
    o = out.cpu().tolist()
    print('  N=%3d: issue %.1f cyc/MMA, retired after %.1f cyc/MMA' % (N, o[0] / n, o[1] / n))
