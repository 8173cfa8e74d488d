"""Dev (GPU box, TF_ATT_TRACE=1 build): where one softmax warp of the attention kernel spends a key block.
Stamps per block: 0 loop top, 1 S ready, 2 S in registers, 3 max done, 4 exps done, 5 PV(j-1) done, 6 P stored, 7 arrived."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
def run(B, NH, T, d, dp):
    q = torch.randn(B * T, NH * dp, device=dev).half(); k = torch.randn(B * T, NH * dp, device=dev).half()
    vt = torch.randn(NH * dp, B * T, device=dev).half(); out = torch.empty(B * T, NH * d, device=dev).half()
    tl = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    for i in range(3):
        b200.tf_attention_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_attention_f16(q.data_ptr(), NH * dp, k.data_ptr(), NH * dp, vt.data_ptr(), B * T, out.data_ptr(), T * NH * d, d, NH * d, B, NH, T, T, T, d, dp, 1 / math.sqrt(d), S()), "attn")
    b200.tf_attention_set_timeline(None)
    torch.cuda.synchronize()
    t = tl.view(64, 8).cpu()
    nb = min(64, (T + 63) // 64)
    t = t[:nb]
    names = ["wait S", "LDTM", "max", "exp+pack", "wait PV", "STS", "fence+arrive", "loop back"]
    print(f"attention B={B} NH={NH} T={T} d={d}: block period (median) {int((t[1:, 0] - t[:-1, 0]).median())} cycles")
    seg = torch.cat([t[:, 1:] - t[:, :-1], torch.cat([t[1:, 0] - t[:-1, 7], torch.zeros(1, dtype=torch.int64)])[:, None]], dim=1)
    med = seg[2:-1].median(dim=0).values
    print("   " + ", ".join(f"{n} {int(v)}" for n, v in zip(names, med)))
run(2, 8, 4096, 40, 48)
run(2, 8, 1024, 80, 80)
