"""GPU probe: where one CTA of the pre-split tensor-core GEMM spends its cycles (per pipeline phase)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine
from fgs_nerf_b200._lib import lib, check
dev = torch.device("cuda:0")
M, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 27000, 256, 256
engine.set_gemm_tmem_accumulate(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); out = torch.zeros(M, N, device=dev)
bias = torch.randn(N, device=dev); bits = torch.zeros(M, 8, dtype=torch.int32, device=dev)
rows = torch.tensor([M], dtype=torch.int32, device=dev)
dbg = torch.zeros(20, dtype=torch.int64, device=dev)
names = ["consumer: chunk wait + drain", "consumer: epilogues", "-", "producer: wait stage free", "producer: split/store/load issue",
         "issuer: wait a_full", "issuer: wait b_full", "issuer: wait r_empty", "issuer: wait free (B prefetch)", "issuer: total"]
for rep in range(3):
    check(lib.fgs_debug_gemm_phases(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.gemm(0, A, B, out, bias=bias, relu=True, rows_dev=rows, tc="w", bits_out=bits)
    e1.record(); torch.cuda.synchronize()
    check(lib.fgs_debug_gemm_phases(None))
print(f"M={M} N={N} K={K}: kernel+split {e0.elapsed_time(e1)*1e3:.1f} us; CTA(0,0) tiles = {(M+127)//128} / 148 CTAs")
d = dbg.cpu().tolist(); tot = max(d[9], 1)
for n, v in zip(names, d):
    print(f"  {n:34s} {v:9d} cyc  {100*v/tot:5.1f}%   {v/1965:.2f} us")
