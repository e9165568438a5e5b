import sys, math, torch
sys.path.insert(0, "whisper-at_b200")
from whisper_at import _lib
L = _lib.lib()
def run(M, N, K, act, res, tc):
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda(); R = torch.randn(M, N, generator=g).cuda() if res else None
    out = torch.empty(M, N, device="cuda")
    _lib.check(L.wat_dbg_gemm(A.data_ptr(), W.data_ptr(), bias.data_ptr(), R.data_ptr() if res else None, out.data_ptr(), M, N, K, act, tc, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
M = 1500 * 64
for name, N, K, act, res in (("fc1-like (gelu, f32 out)", 5120, 1280, 1, False), ("out-proj (f32 residual)", 1280, 1280, 0, True), ("fc2 (f32 residual)", 1280, 5120, 0, True)):
    print("=====", name, file=sys.stderr); sys.stderr.flush()
    run(M, N, K, act, res, 2); run(M, N, K, act, res, 3)
print("===== fc1 bf16+gelu (16 epilogue warps)", file=sys.stderr); sys.stderr.flush()
run(M, 5120, 1280, 1, False, 4)
