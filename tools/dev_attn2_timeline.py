"""Dev (GPU box, TF_ATT_TRACE=1 build): per-block clock stamps of tf_attention2_kernel, CTA (0,0,0).
softmax warpgroup t (first lane): 0 loop top, 1 S ready, 2 S in registers, 3 max done, 4 turn granted, 5 exps done, 6 PV(j-1) seen, 7 P stored + arrived
MMA thread: 0 loop top, 1 QK(j+1) issued, 2 P0 ready, 3 PV0 issued, 4 P1 ready, 5 PV1 issued"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinyfusers_b200.native.b200.ops import b200
from tinyfusers_b200.attention.attention import _pad64
dev = torch.device("cuda:0"); b200.init(0)
S = lambda: torch.cuda.current_stream().cuda_stream
def run(B, NH, T, d, emu, ones=1):
    dp = (d + 15) // 16 * 16; dvp = _pad64(d)
    Q = torch.randn(B, T, NH, dp, device=dev).half(); K = torch.randn(B, T, NH, dp, device=dev).half()
    V = torch.zeros(B, T, NH, dvp, device=dev).half(); V[..., :d] = torch.randn(B, T, NH, d, device=dev).half()
    if ones: V[..., d] = 1.0
    out = torch.empty(B, T, NH, d, device=dev).half()
    tl = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
    b200.tf_attention_set_variant(2, emu)
    for i in range(3):
        b200.tf_attention_set_timeline(tl.data_ptr() if i == 2 else None)
        b200.check(b200.tf_attention_v_f16(Q.data_ptr(), NH * dp, K.data_ptr(), NH * dp, V.data_ptr(), NH * dvp, out.data_ptr(), T * NH * d, d, NH * d,
                                           B, NH, T, T, T, d, dp, dvp, 1 / math.sqrt(d), 2 if ones else 0, S()), "attn")
    b200.tf_attention_set_timeline(None); b200.tf_attention_set_variant(0, -1)
    torch.cuda.synchronize()
    t = tl.view(3, 64, 8).cpu()
    nb = min(64, T // 64)
    print(f"attention2 B={B} NH={NH} T={T} d={d} emu={emu} ones={ones}")
    for who, names in ((0, ["wait S", "LDTM", "max", "wait turn", "exp+pack", "wait PV", "STTM+arrive", "loop back"]),
                       (1, ["wait S", "LDTM", "max", "wait turn", "exp+pack", "wait PV", "STTM+arrive", "loop back"]),
                       (2, ["issue QK", "wait P0", "issue PV0", "wait P1", "issue PV1", "loop back"])):
        w = t[who, :nb]
        n = len(names)
        cols = w[:, :n] if who < 2 else w[:, :6]
        nxt = torch.cat([w[1:, 0], w[-1:, 0]])
        seg = torch.cat([cols[:, 1:] - cols[:, :-1], (nxt - cols[:, -1])[:, None]], dim=1)
        med = seg[4:-2].median(dim=0).values
        print(f"  {'softmax' + str(who) if who < 2 else 'mma    '}: period {int((w[5:-1, 0] - w[4:-2, 0]).median())}  " + ", ".join(f"{nm} {int(v)}" for nm, v in zip(names, med)))
    # relative phase of the two softmax warpgroups and the mma thread in one steady-state block
    j = nb // 2
    base = int(t[0, j, 0])
    print("  block", j, "softmax0", [int(x) - base for x in t[0, j]], "softmax1", [int(x) - base for x in t[1, j]], "mma", [int(x) - base for x in t[2, j, :6]])
for emu, ones in ((0, 0), (0, 1), (2, 1), (4, 1)):
    run(2, 8, 4096, 40, emu, ones)
