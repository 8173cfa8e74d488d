"""Focused profiling target (GPU box, under ncu): a few representative GEMM/conv/attention launches."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
ws = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def conv(NI, H, W, Cin, Cout, reps=3):
    x = torch.randn(NI, H, W, Cin, device=dev).half(); w = (torch.randn(Cout, 3, 3, Cin, device=dev) / 50).half()
    b = torch.randn(Cout, device=dev); out = torch.empty(NI, H, W, Cout, dtype=torch.half, device=dev)
    for _ in range(reps):
        b200.check(b200.tf_conv2d_nhwc_f16(x.data_ptr(), NI, H, W, Cin, Cin, w.data_ptr(), Cout, 3, 1, out.data_ptr(), Cout, b.data_ptr(), None, 0, 0, ws.data_ptr(), ws.numel(), S()), "conv")
def gemm(M, N, K, geglu=False, reps=3):
    A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) / 30).half(); b = torch.randn(N, device=dev)
    No = N // 2 if geglu else N
    out = torch.empty(M, No, dtype=torch.half, device=dev)
    for _ in range(reps):
        b200.check(b200.tf_gemm_f16(A.data_ptr(), K, W.data_ptr(), K, out.data_ptr(), No, M, N, K, b.data_ptr(), None, 0, 2 if geglu else 0, ws.data_ptr(), ws.numel(), S()), "gemm")
def attn(B, NH, T, d, reps=3):
    """the production path: V in its natural layout, head dim padded like the model pads it (tf_attention_v_f16)"""
    from tinyfusers_b200.attention.attention import _pad64
    dp = (d + 15) // 16 * 16; dvp = _pad64(d)
    Q = torch.zeros(B, T, NH, dp, device=dev).half(); Q[..., :d] = torch.randn(B, T, NH, d, device=dev)
    K = torch.zeros(B, T, NH, dp, device=dev).half(); K[..., :d] = torch.randn(B, T, NH, d, device=dev)
    V = torch.zeros(B, T, NH, dvp, device=dev).half(); V[..., :d] = torch.randn(B, T, NH, d, device=dev)
    out = torch.empty(B, T, NH, d, dtype=torch.half, device=dev)
    for _ in range(reps):
        b200.check(b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(), T * NH * d, d, NH * d,
                                           B, NH, T, T, T, d, dp, dvp, 1 / math.sqrt(d), 0, S()), "attn")
conv(2, 64, 64, 320, 320)
conv(2, 32, 32, 640, 640)
gemm(8192, 2560, 320)
gemm(8192, 320, 320)
attn(2, 8, 4096, 40)
conv(1, 512, 512, 128, 128, reps=2)      # VAE decoder, last stage (M = 262 144 pixels)
attn(16, 8, 4096, 40, reps=2)            # C4 (8 images / GPU): the two-tile kernel
attn(2, 8, 1024, 80)
torch.cuda.synchronize()
print("done")
