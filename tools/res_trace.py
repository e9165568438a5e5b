"""Clock trace of the residual-producer GEMM's first epilogue warp (test hook wat_dbg_ln_gemm, WAT_DBG_TRACE=1) at the
out-proj and fc2 shapes of large-v2, 64 clips.  Run on a GPU box:  WAT_DBG_TRACE=1 python tools/res_trace.py"""
import math, sys, torch
sys.path.insert(0, "whisper-at_b200")
from whisper_at import _lib
L = _lib.lib()
def run(M, D, K1, N2):
    g = torch.Generator().manual_seed(0)
    A1 = (torch.randn(M, K1, generator=g)).to(torch.bfloat16).cuda(); W1 = (torch.randn(D, K1, generator=g) / math.sqrt(K1)).to(torch.bfloat16).cuda()
    b1 = torch.randn(D, generator=g).cuda(); R = torch.randn(M, D, generator=g).to(torch.float16).cuda()
    x = torch.empty(M, D, dtype=torch.float16, device="cuda"); xb = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    stats = torch.empty(M, 16, 2, device="cuda")
    W2 = (torch.randn(N2, D, generator=g) / math.sqrt(D)).cuda(); gam = torch.ones(D).cuda(); bet = torch.zeros(D).cuda(); b2 = torch.zeros(N2).cuda()
    out = torch.empty(M, N2, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        _lib.check(L.wat_dbg_ln_gemm(A1.data_ptr(), W1.data_ptr(), b1.data_ptr(), R.data_ptr(), x.data_ptr(), xb.data_ptr(), stats.data_ptr(), None,
                                     W2.data_ptr(), gam.data_ptr(), bet.data_ptr(), b2.data_ptr(), out.data_ptr(), M, D, K1, N2, 0, 1, 1,
                                     torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
M = 1500 * 64
for name, K1 in (("out-proj K=1280", 1280), ("fc2 K=5120", 5120)):
    print("=====", name, file=sys.stderr); sys.stderr.flush()
    run(M, 1280, K1, 3840)
