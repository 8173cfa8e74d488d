"""The cond / uncond halves of classifier-free guidance on two GPUs: one image, lower latency per step.

Reference: `StableDiffusion.get_model_output` (tinyfusers/variants/sd.py:27-46) runs the UNet on the batch [uncond ; cond] and
combines `e_t = u + g (c - u)`; the halves are independent until that line ("latency mode"). Here, with two
processes (one per GPU, torch.distributed world size 2):

  rank 0: UNet(latent, uncond context) at batch 1      rank 1: UNet(latent, cond context) at batch 1
  each: ONE kernel (csrc/tf_p2p.cu) stores its eps (64 KiB at 512^2) into the peer's mailbox over NVLink - peer-mapped memory,
        plain stores + a sequence flag, no NCCL launch, no host round trip - waits for the peer's flag, and applies CFG + DDIM to
        its own copy of the latent. Same expression, same operands, same rounding on both GPUs: the latents stay bit-identical.

The whole step (UNet + exchange/update + counters) is one captured CUDA graph per rank, replayed `len(timesteps)` times.
torch.distributed is used once, at set-up, to swap the 64-byte CUDA IPC handles of the mailboxes, and for a barrier before the
loop (the exchange kernel spins on the peer's flag, so the two ranks must enter the loop together)."""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import get_layernorm_strided, get_quirks, packing
from .native.b200.ops import b200
from .runtime import F32, stream_ptr


class CfgSplitSampler:
    """Device-resident sampler state of ONE image on this rank's GPU; rank 0 owns the unconditional half."""

    def __init__(self, unet, H, W, ctx_tokens, group=None):
        if not dist.is_initialized() or dist.get_world_size(group) != 2:
            raise RuntimeError("CfgSplitSampler needs an initialised torch.distributed group of exactly 2 ranks (one per GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.H, self.W = H, W
        self.engine = unet.engine(1, H, W, n_src=1, ctx_tokens=ctx_tokens)
        e = self.engine
        dev = e.latent.device
        self.latent, self.context = e.latent, e.context            # (1,4,H,W) fp32; (1,T,768) fp32: this rank's half
        self.e_t = torch.zeros((1, 4, H, W), dtype=F32, device=dev)
        self.tab = torch.ones((3, 1024), dtype=F32, device=dev)
        self.tab[0].zero_()
        self.idx = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seq = torch.zeros(1, dtype=torch.int32, device=dev)
        self.guidance = 7.5
        self._graphs, self._graphs_gen = {}, None
        # ---- mailbox + flags: cudaMalloc'ed (CUDA IPC names whole allocations), zeroed, mapped into the peer ----
        total = 4 * H * W
        self.nblk = int(b200.tf_p2p_blocks(4, H * W))
        self.mail_bytes = 2 * total * 4
        nbytes = self.mail_bytes + 2 * self.nblk * 4
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        b200.check(b200.tf_p2p_alloc(nbytes, ctypes.byref(ptr), handle), "tf_p2p_alloc")
        self.my_buf = ptr.value
        handles = [None, None]
        dist.all_gather_object(handles, bytes(handle), group=group)
        peer = (ctypes.c_ubyte * 64).from_buffer_copy(handles[1 - self.rank])
        pptr = ctypes.c_void_p()
        b200.check(b200.tf_p2p_open(peer, ctypes.byref(pptr)), "tf_p2p_open")
        self.peer_buf = pptr.value
        dist.barrier(group=group)

    def close(self):
        if getattr(self, "peer_buf", None):
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            b200.tf_p2p_close(self.peer_buf)
            b200.tf_p2p_free(self.my_buf)
            self.peer_buf = self.my_buf = None

    def load(self, unconditional_context, context, latent):
        mine = unconditional_context if self.rank == 0 else context
        self.latent.copy_(latent.reshape(self.latent.shape), non_blocking=True)
        self.context.copy_(mine.reshape(self.context.shape), non_blocking=True)

    def set_tables(self, timesteps, alphas, alphas_prev, guidance):
        n = len(timesteps)
        self.tab[0, :n].copy_(torch.as_tensor(np.asarray(timesteps, dtype=np.float32)))
        self.tab[1, :n].copy_(torch.as_tensor(alphas, dtype=F32).reshape(-1)[:n])
        self.tab[2, :n].copy_(torch.as_tensor(alphas_prev, dtype=F32).reshape(-1)[:n])
        self.idx.fill_(n - 1)
        self.guidance = float(guidance)

    def enqueue_step(self):
        """UNet (this rank's half) -> exchange + CFG + DDIM in place -> idx -= 1, seq += 1; all on the current stream."""
        e = self.engine
        e._enqueue(t_ptr=self.tab[0].data_ptr(), idx_ptr=self.idx.data_ptr())
        flags_off = self.mail_bytes
        st = b200.tf_cfg_ddim_step_split_f32(e.eps.data_ptr(), 16, self.latent.data_ptr(), self.latent.data_ptr(),
                                             self.e_t.data_ptr(), self.tab[1].data_ptr(), self.tab[2].data_ptr(),
                                             self.idx.data_ptr(), float(self.guidance), 4, self.H * self.W, self.rank,
                                             self.my_buf, self.peer_buf, self.my_buf + flags_off, self.peer_buf + flags_off,
                                             self.seq.data_ptr(), 3, stream_ptr())
        b200.check(st, "tf_cfg_ddim_step_split_f32")
        b200.check(b200.tf_add_int(self.idx.data_ptr(), -1, stream_ptr()), "tf_add_int")
        b200.check(b200.tf_add_int(self.seq.data_ptr(), 1, stream_ptr()), "tf_add_int")

    def _graph(self):
        gen = packing.generation("unet")
        if self._graphs_gen != gen:
            self._graphs = {}
        g = self._graphs.get(self.guidance)
        if g is None:
            # eager warm-up step (weight packing never inside a capture); BOTH ranks run it, so the exchange pairs up
            lat, idx = self.latent.clone(), self.idx.clone()
            self.enqueue_step()
            torch.cuda.synchronize()
            self.latent.copy_(lat)
            self.idx.copy_(idx)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue_step()
            self.latent.copy_(lat)
            self.idx.copy_(idx)
            # the warm-up advanced seq by one on both ranks (kept: it only has to agree between the ranks); the capture itself
            # executes nothing
            self._graphs[self.guidance] = g
            self._graphs_gen = packing.generation("unet")
        return g

    def run(self, n_steps, use_graph=True):
        g = self._graph() if use_graph else None
        torch.cuda.synchronize()
        dist.barrier(group=self.group)            # enter the loop together: the exchange kernel spins on the peer
        for _ in range(n_steps):
            if g is not None:
                g.replay()
            else:
                self.enqueue_step()


def sample_cfg_split(model, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance, group=None,
                     use_graph=True):
    """One image on two GPUs (see the module docstring). Both ranks pass the same arguments and both return the final latent."""
    B, _, H, W = latent.shape
    if B != 1:
        raise RuntimeError("sample_cfg_split denoises ONE image on two GPUs; shard larger batches by image (sample_dp)")
    key = ("cfg_split", torch.cuda.current_device(), H, W, context.shape[1], get_quirks(), get_layernorm_strided())
    s = model._samplers.get(key)
    if s is None:
        s = model._samplers[key] = CfgSplitSampler(model.model.diffusion_model, H, W, context.shape[1], group)
    s.load(unconditional_context, context, latent)
    s.set_tables(timesteps, alphas, alphas_prev, guidance)
    s.run(len(timesteps), use_graph=use_graph)
    return s.latent.clone()
