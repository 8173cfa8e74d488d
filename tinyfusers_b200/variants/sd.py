"""StableDiffusion sampler object with the reference's signatures (reference: tinyfusers/variants/sd.py:7-65).

One sampler step = UNet forward at batch 2B ([uncond ; cond]) + CFG combine + DDIM (eta = 0) update.
The reference round-trips the latent batch through host memory and synchronises the device twice per
step (sd.py:34-41); here the step is a fixed launch sequence on one stream — conv_in forms the CFG batch
by index, one fused kernel does CFG + DDIM in place — and `sample()` captures it once into a CUDA graph
and replays it with a device-resident step counter selecting (t, a_t, a_prev) from device tables.

`decode` (post_quant_conv -> VAE decoder -> uint8 image) and `cond_stage_model` (CLIP text encoder) are the rows next
to the hot path; they run on the same kernels (tinyfusers_b200/vae).
"""
from collections import namedtuple

import numpy as np
import torch

from .. import fp32, get_layernorm_strided, get_quirks, packing, weight_prefetch_enabled
from ..native.b200.ops import b200
from ..runtime import F32, require_cuda, stream_ptr
from ..vae.encoder import CLIPTextTransformer
from ..vae.vae import AutoencoderKL
from ..vision.unet import UNetModel


def get_alphas_cumprod(beta_start=0.00085, beta_end=0.0120, n_training_steps=1000):
    """fp32 scaled-linear beta schedule, computed once on the host (reference: sd.py:61-65)."""
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n_training_steps, dtype=np.float32) ** 2
    alphas = (1.0 - betas).astype(np.float32)
    ac = torch.from_numpy(np.cumprod(alphas, axis=0).astype(np.float32))
    return ac.cuda() if torch.cuda.is_available() else ac


class StableDiffusion:
    def __init__(self):
        self.alphas_cumprod = get_alphas_cumprod()
        self.model = namedtuple("DiffusionModel", ["diffusion_model"])(diffusion_model=UNetModel())
        self.first_stage_model = AutoencoderKL()
        self.cond_stage_model = namedtuple("CondStageModel", ["transformer"])(
            transformer=namedtuple("Transformer", ["text_model"])(text_model=CLIPTextTransformer()))
        self._samplers = {}

    # ---- reference API -------------------------------------------------------------------------
    def get_x_prev_and_pred_x0(self, x, e_t, a_t, a_prev):
        require_cuda(x, "x")
        x32, e32 = x.to(F32).contiguous(), e_t.to(F32).contiguous()
        a_t = torch.as_tensor(a_t, dtype=F32, device=x.device).reshape(-1)[:1].contiguous()
        a_prev = torch.as_tensor(a_prev, dtype=F32, device=x.device).reshape(-1)[:1].contiguous()
        x_prev, pred_x0 = torch.empty_like(x32), torch.empty_like(x32)
        st = b200.tf_ddim_step_f32(x32.data_ptr(), e32.data_ptr(), a_t.data_ptr(), a_prev.data_ptr(), x_prev.data_ptr(),
                                   pred_x0.data_ptr(), x32.numel(), stream_ptr())
        b200.check(st, "tf_ddim_step_f32")
        return x_prev, pred_x0

    def get_model_output(self, unconditional_context, context, latent, timestep, unconditional_guidance_scale):
        if fp32.enabled():
            return fp32.get_model_output(self, unconditional_context, context, latent, timestep, unconditional_guidance_scale)
        s = self._sampler(latent.shape, context.shape[1])
        s.load(unconditional_context, context, latent)
        s.set_scalars(timestep, 1.0, 1.0, unconditional_guidance_scale)
        s.enqueue_step(update_latent=False)
        return s.e_t.clone()

    def decode(self, x):
        """Latent (B,4,h,w) -> uint8 image (8h, 8w, 3) for B = 1 (as the reference), (B, 8h, 8w, 3) otherwise
        (reference: sd.py:48-54, whose reshape(3,512,512) is generalised to the decoded size)."""
        require_cuda(x, "x")
        if fp32.enabled():
            return fp32.decode(self, x)
        fsm = self.first_stage_model
        z = fsm.post_quant(x, 1 / 0.18215)
        eng = fsm.decoder._engine(tuple(z.shape))
        img = eng.decode_nhwc_f32(z)                       # (B, H, W, 8) fp32 NHWC, channels 0..2 valid
        B, H, W, _ = img.shape
        out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=x.device)
        st = b200.tf_image_to_u8(img.data_ptr(), 8, out.data_ptr(), B * H * W, 3, stream_ptr())
        b200.check(st, "tf_image_to_u8")
        return out[0] if B == 1 else out

    def __call__(self, unconditional_context, context, latent, timestep, alphas, alphas_prev, guidance):
        if fp32.enabled():      # the reference's own two calls (sd.py:56-59), eager, fp32 kernels
            e_t = self.get_model_output(unconditional_context, context, latent, timestep, guidance)
            return self.get_x_prev_and_pred_x0(latent, e_t, alphas, alphas_prev)[0]
        s = self._sampler(latent.shape, context.shape[1])
        s.load(unconditional_context, context, latent)
        s.set_scalars(timestep, alphas, alphas_prev, guidance)
        s.step_once()
        return s.latent.clone()

    # ---- whole sampler loop on device (graph-captured steps) -----------------------------------------
    def sample(self, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance, use_graph=True):
        """Runs len(timesteps) DDIM steps, last-to-first like example/sd1.py:68-73, and returns the final latent."""
        if fp32.enabled():      # parity mode: the reference's own host loop (example/sd1.py:68-73) over __call__, eager
            a = torch.as_tensor(alphas, dtype=F32).reshape(-1)
            ap = torch.as_tensor(alphas_prev, dtype=F32).reshape(-1)
            for i in reversed(range(len(timesteps))):
                latent = self(unconditional_context, context, latent, torch.tensor([float(timesteps[i])]),
                              a[i:i + 1], ap[i:i + 1], guidance)
            return latent
        s = self._sampler(latent.shape, context.shape[1])
        s.load(unconditional_context, context, latent)
        s.set_tables(timesteps, alphas, alphas_prev, guidance)
        s.run(len(timesteps), use_graph=use_graph)
        return s.latent.clone()

    def sample_dp(self, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance, group=None):
        """`sample` sharded by image over the ranks of an initialised torch.distributed group (one process per GPU, one
        UNet replica each; the only collective is the all-gather of the final latents, tinyfusers_b200/dp.py). Every rank
        passes the full batch and receives all final latents. Reference: the loop of example/sd1.py:68-73 per image."""
        from .. import dp
        return dp.sample_sharded(self, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance, group)

    def sample_cfg_split(self, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance, group=None):
        """ONE image on TWO GPUs: rank 0 evaluates the unconditional half of `get_model_output`'s batch (reference sd.py:27-46),
        rank 1 the conditional half; the two noise predictions are exchanged by peer stores over NVLink inside the CFG + DDIM
        kernel (tinyfusers_b200/cfg_split.py). Both ranks return the same final latent, bit for bit."""
        from .. import cfg_split
        return cfg_split.sample_cfg_split(self, unconditional_context, context, latent, timesteps, alphas, alphas_prev, guidance,
                                          group)

    def _sampler(self, latent_shape, ctx_tokens):
        B, C, H, W = latent_shape
        key = (torch.cuda.current_device(), B, H, W, ctx_tokens, get_quirks(), get_layernorm_strided())
        s = self._samplers.get(key)
        if s is None:
            s = SamplerEngine(self.model.diffusion_model, B, H, W, ctx_tokens)
            self._samplers[key] = s
        return s


def _scalar(v):
    if isinstance(v, torch.Tensor):
        return float(v.reshape(-1)[0].item())
    return float(np.asarray(v).reshape(-1)[0])


class SamplerEngine:
    """Device-resident sampler state for B images: latent, [uncond ; cond] context, schedule tables, step index."""

    def __init__(self, unet, B, H, W, ctx_tokens):
        self.B, self.H, self.W = B, H, W
        self.unet_engine = unet.engine(2 * B, H, W, n_src=B, ctx_tokens=ctx_tokens)
        e = self.unet_engine
        dev = e.latent.device
        self.latent = e.latent                      # (B,4,H,W) fp32, updated in place every step
        self.context = e.context                    # (2B,T,768) fp32: [uncond ; cond]  (sd.py:32)
        self.e_t = torch.zeros((B, 4, H, W), dtype=F32, device=dev)
        self.max_steps = 1024
        # schedule tables: one (3, max_steps) buffer = [t ; a_t ; a_prev], so that the per-call scalars of the drop-in
        # __call__ (column 0 of each row) arrive in ONE host->device copy from a pinned staging buffer
        self.tab = torch.ones((3, self.max_steps), dtype=F32, device=dev)
        self.tab[0].zero_()
        self.t_tab, self.a_tab, self.ap_tab = self.tab[0], self.tab[1], self.tab[2]
        self._scal_host = torch.zeros(3, dtype=F32).pin_memory()
        self._scal_last = None
        self._idx_zero = True
        self.idx = torch.zeros(1, dtype=torch.int32, device=dev)
        self.guidance = 7.5

    def load(self, unconditional_context, context, latent):
        """Device tensors, or host tensors (pinned: asynchronous) copied straight into the sampler's resident buffers."""
        B = self.B
        self.latent.copy_(latent.reshape(self.latent.shape), non_blocking=True)
        self.context[:B].copy_(unconditional_context.reshape(B, -1, 768), non_blocking=True)
        self.context[B:].copy_(context.reshape(B, -1, 768), non_blocking=True)

    def set_scalars(self, timestep, a_t, a_prev, guidance):
        vals = (_scalar(timestep), _scalar(a_t), _scalar(a_prev))
        if vals != self._scal_last:
            # the pinned staging buffer may still be the source of the previous call's copy
            ev = getattr(self, "_scal_event", None)
            if ev is not None:
                ev.synchronize()
            self._scal_host[0], self._scal_host[1], self._scal_host[2] = vals
            self.tab[:, 0].copy_(self._scal_host, non_blocking=True)
            self._scal_event = torch.cuda.Event()
            self._scal_event.record()
            self._scal_last = vals
        if not self._idx_zero:
            self.idx.zero_()
            self._idx_zero = True
        self.guidance = _scalar(guidance)

    def set_tables(self, timesteps, alphas, alphas_prev, guidance):
        n = len(timesteps)
        assert n <= self.max_steps
        self.t_tab[:n].copy_(torch.as_tensor(np.asarray(timesteps, dtype=np.float32)))
        self.a_tab[:n].copy_(torch.as_tensor(alphas, dtype=F32).reshape(-1)[:n])
        self.ap_tab[:n].copy_(torch.as_tensor(alphas_prev, dtype=F32).reshape(-1)[:n])
        self.idx.fill_(n - 1)   # the loop runs from the last timestep to the first
        self._idx_zero, self._scal_last = False, None
        self.guidance = _scalar(guidance)

    def enqueue_step(self, update_latent=True, advance=False):
        """One step on the current stream: UNet(2B) -> CFG + DDIM (in place) [-> idx -= 1]."""
        e = self.unet_engine
        e._enqueue(t_ptr=self.t_tab.data_ptr(), idx_ptr=self.idx.data_ptr())
        if e.ctx.skip("misc"):
            return
        # when only e_t is wanted (get_model_output) the latent is left untouched
        dst = self.latent.data_ptr() if update_latent else self._scratch().data_ptr()
        st = b200.tf_cfg_ddim_step_f32(e.eps.data_ptr(), 16, self.latent.data_ptr(), dst, self.e_t.data_ptr(),
                                       self.a_tab.data_ptr(), self.ap_tab.data_ptr(), self.idx.data_ptr(),
                                       float(self.guidance), self.B, 4, self.H * self.W, stream_ptr())
        b200.check(st, "tf_cfg_ddim_step_f32")
        if advance:
            b200.check(b200.tf_add_int(self.idx.data_ptr(), -1, stream_ptr()), "tf_add_int")
            self._idx_zero = False

    def _scratch(self):
        if not hasattr(self, "_scratch_buf"):
            self._scratch_buf = torch.empty_like(self.latent)
        return self._scratch_buf

    def capture(self, advance=True):
        """Capture one step into a CUDA graph; replays need no host work. advance=True also decrements the
        device step index (sampler loop); advance=False is the graph behind the drop-in __call__."""
        pf = weight_prefetch_enabled()
        if pf or not getattr(self, "_warm", False):
            # eager pass: lazy attribute setup and weight packing (never inside a capture) - and, for the next-layer weight
            # prefetch, the library records this step's weight sequence (tf_weight_prefetch_mode 1) so that the captured
            # launches below can each name the weights of the launch behind them (mode 2)
            lat, idx = self.latent.clone(), self.idx.clone()
            if pf and getattr(self, "_warm", False):
                b200.check(b200.tf_weight_prefetch_mode(1), "tf_weight_prefetch_mode")
            self.enqueue_step(advance=False)
            torch.cuda.synchronize()
            if pf and not getattr(self, "_warm", False):      # the first pass packed weights: record a settled second one
                b200.check(b200.tf_weight_prefetch_mode(1), "tf_weight_prefetch_mode")
                self.enqueue_step(advance=False)
                torch.cuda.synchronize()
            self.latent.copy_(lat)
            self.idx.copy_(idx)
            self._warm = True
        lat = self.latent.clone()
        g = torch.cuda.CUDAGraph()
        try:
            if pf:
                b200.check(b200.tf_weight_prefetch_mode(2), "tf_weight_prefetch_mode")
            with torch.cuda.graph(g):
                self.enqueue_step(update_latent=True, advance=advance)
        finally:
            b200.tf_weight_prefetch_mode(0)
        self.latent.copy_(lat)
        return g

    def _graph(self, advance):
        """The captured step for (advance, guidance), valid for the CURRENT weights: a graph bakes in the addresses of the
        packed fp16 weights, so when `packing.generation()` has moved since capture (update_state, a weight assigned and
        repacked by an eager call) every graph of this engine is dropped and the step is re-captured after an eager
        warm-up that repacks - a stale graph would replay old (or freed) weights."""
        gen = packing.generation("unet")
        if getattr(self, "_graphs_gen", None) != gen:
            self._graphs, self._warm = {}, False
        key = (advance, self.guidance, weight_prefetch_enabled())
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = self.capture(advance)
            # the warm-up step inside capture() may have repacked (bumping the generation): the graphs are valid for the
            # generation seen AFTER it, provided nothing moved during the capture itself (packing.cached raises there)
            self._graphs_gen = packing.generation("unet")
        return g

    def step_once(self, use_graph=True):
        if use_graph:
            self._graph(False).replay()
        else:
            self.enqueue_step(update_latent=True, advance=False)

    def run(self, n_steps, use_graph=True):
        self._idx_zero = False
        if use_graph:
            g = self._graph(True)
            for _ in range(n_steps):
                g.replay()
        else:
            for _ in range(n_steps):
                self.enqueue_step(update_latent=True, advance=True)
