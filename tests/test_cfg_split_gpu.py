"""CFG halves on two GPUs (tinyfusers_b200/cfg_split.py, csrc/tf_p2p.cu): the exchange + CFG + DDIM kernel on ONE GPU, the two
ranks driven one after the other (rank 1 send-only, then rank 0 send + wait + update; then the mirror image) through buffers
from tf_p2p_alloc - against the fused single-GPU kernel tf_cfg_ddim_step_f32 on the [uncond ; cond] batch. Bit-exact: both
evaluate the same fp32 expression. (Kernels that spin on each other are never co-scheduled on one GPU - B200_PROFILING.md; the
true two-process run is tools/run_cfg_split.py under `gpurun --gpus 2`.)"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_split_kernel_matches_fused_step_bit_for_bit():
    from tinyfusers_b200.native.b200.ops import b200
    b200.init(0)
    S = torch.cuda.current_stream().cuda_stream
    H = W = 64
    HW, C = H * W, 4
    g = torch.Generator().manual_seed(3)
    eps = torch.randn(2, HW, 16, generator=g).cuda()                  # fp32 NHWC, stride 16: [uncond ; cond]
    lat = torch.randn(1, C, H, W, generator=g).cuda()
    a_tab = torch.tensor([0.9, 0.5, 0.3]).cuda()
    ap_tab = torch.tensor([1.0, 0.9, 0.5]).cuda()
    idx = torch.tensor([1], dtype=torch.int32).cuda()
    ref_lat, ref_e = torch.empty_like(lat), torch.empty_like(lat)
    b200.check(b200.tf_cfg_ddim_step_f32(eps.data_ptr(), 16, lat.data_ptr(), ref_lat.data_ptr(), ref_e.data_ptr(), a_tab.data_ptr(),
                                         ap_tab.data_ptr(), idx.data_ptr(), 7.5, 1, C, HW, S), "fused")
    nblk = b200.tf_p2p_blocks(C, HW)
    mail = 2 * C * HW * 4
    bufs = []
    for _ in range(2):                                                # one mailbox + flags per emulated rank
        p, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        b200.check(b200.tf_p2p_alloc(mail + 2 * nblk * 4, ctypes.byref(p), h), "tf_p2p_alloc")
        bufs.append(p.value)
    try:
        for seq_val in (0, 1, 2):                                     # both slot parities, flags reused across steps
            seq = torch.tensor([seq_val], dtype=torch.int32).cuda()
            outs = []
            for first in (1, 0):                                      # which rank only sends; the other then completes
                other = 1 - first
                lat_o, e_o = torch.empty_like(lat), torch.empty_like(lat)

                def call(rank, mode, lo, eo):
                    st = b200.tf_cfg_ddim_step_split_f32(eps[rank].data_ptr(), 16, lat.data_ptr(), lo.data_ptr(), eo.data_ptr(),
                                                         a_tab.data_ptr(), ap_tab.data_ptr(), idx.data_ptr(), 7.5, C, HW, rank,
                                                         bufs[rank], bufs[1 - rank], bufs[rank] + mail, bufs[1 - rank] + mail,
                                                         seq.data_ptr(), mode, S)
                    b200.check(st, "split")
                call(first, 1, lat_o, e_o)                            # send only: fills the other rank's mailbox and flags
                call(other, 3, lat_o, e_o)                            # send + wait (already satisfied) + update
                torch.cuda.synchronize()
                outs.append((lat_o, e_o))
            for lat_o, e_o in outs:
                assert torch.equal(lat_o, ref_lat) and torch.equal(e_o, ref_e), seq_val
    finally:
        for p in bufs:
            b200.tf_p2p_free(p)
