"""GPU probe: fgs_gemm_tn (one dW problem through mlp_dw_kernel) timing + CTA-0 phases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine
from fgs_nerf_b200._lib import lib, check, ptr, stream
dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 27000
M = N = 256
A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev)
out = torch.zeros(M, N, device=dev); cs = torch.zeros(M, device=dev)
rows = torch.tensor([K], dtype=torch.int32, device=dev)
dbg = torch.zeros(20, dtype=torch.int64, device=dev)
names = ["consumer: wait+drain", "consumer: epilogue", "-", "producer: wait free", "producer: work", "issuer: wait a_full", "-",
         "issuer: wait r_empty", "slabs of CTA 0", "issuer: total"]
for mode in (0, 0):
  for rep in range(2):
    check(lib.fgs_debug_gemm_phases(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.fgs_gemm_tn(M, N, K, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(out), out.stride(0), ptr(cs), ptr(rows), stream()))
    e1.record(); torch.cuda.synchronize()
    check(lib.fgs_debug_gemm_phases(None))
  d = dbg.cpu().tolist()[10:]
  print(f"{e0.elapsed_time(e1)*1e3:.1f} us; CTA0 slabs {d[8]}: producer work {d[4]/max(d[8],1):.0f} cyc/slab, wait free {d[3]/max(d[8],1):.0f}; issuer total {d[9]/max(d[8],1):.0f} cyc/slab, wait a_full {d[5]/max(d[8],1):.0f}, wait r_empty {d[7]/max(d[8],1):.0f}; consumer drain {d[0]/max(d[8],1):.0f} epi {d[1]}")
