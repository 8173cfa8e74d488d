"""Dev (GPU box): a handful of plain (not graph-captured) launches of one self-attention shape, for `ncu --set full`.
argv: version emu ones [B NH T d]"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.attention.attention import _pad64
ver, emu, ones = (int(x) for x in sys.argv[1:4])
B, NH, T, d = (int(x) for x in sys.argv[4:8]) if len(sys.argv) >= 8 else (2, 8, 4096, 40)
dev = torch.device("cuda:0"); b200.init(0)
dp = (d + 15) // 16 * 16; dvp = _pad64(d)
Q = torch.zeros(B, T, NH, dp, dtype=torch.half, device=dev); Q[..., :d] = torch.randn(B, T, NH, d, device=dev)
K = torch.zeros(B, T, NH, dp, dtype=torch.half, device=dev); K[..., :d] = torch.randn(B, T, NH, d, device=dev)
V = torch.zeros(B, T, NH, dvp, dtype=torch.half, device=dev); V[..., :d] = torch.randn(B, T, NH, d, device=dev)
if ones: V[..., d] = 1.0
out = torch.zeros(B, T, NH, d, dtype=torch.half, device=dev)
b200.check(b200.tf_attention_set_variant(ver, emu), "variant")
for _ in range(6):
    b200.check(b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(), T * NH * d, d, NH * d,
                                       B, NH, T, T, T, d, dp, dvp, 1 / math.sqrt(d), 2 if ones else 0, torch.cuda.current_stream().cuda_stream), "attn")
torch.cuda.synchronize()
print("ok", float(out.float().abs().max()))
