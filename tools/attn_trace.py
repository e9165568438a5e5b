import sys, math, torch
sys.path.insert(0, "whisper-at_b200")
from whisper_at import _lib
B, T, H = 8, 1500, 20
D = 64 * H
g = torch.Generator().manual_seed(0)
x = torch.randn(B * T, D, generator=g).cuda()
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
w = (torch.randn(3 * D, D, generator=g) / math.sqrt(D) * scale).cuda()
b = torch.randn(3 * D, generator=g).cuda()
out = torch.empty(B * T, D, device="cuda")
for tc in (1, 3):
    _lib.check(_lib.lib().wat_dbg_attention(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, T, H, tc, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
