"""LIBRARY BASELINE (measurement infrastructure, never on the product path): the reference's UNet CFG step through the
vendor libraries the reference itself calls - cuDNN `conv_fprop` graphs built with the cuDNN python frontend exactly as
`/root/reference/tinyfusers/vision/conv2d.py:9-28` (fp32) and `:31-46` (fp16 I/O, fp32 compute) declare them, cuBLAS GEMMs
for every Linear and for QK^T / PV (`ff/linear.py:119-120`, `attention/sdpa.py:66,76`; torch.matmul is the same cuBLAS call
CuPy makes), library elementwise / normalisation kernels for the rest - on the same B200, same seeded weights and inputs,
CUDA-graph captured. BASELINE.md section 3 names three variants; `bench.py` reports them in its `library_baseline` block:

  R-cached   fp32 as the reference (NCHW in, cuDNN's NHWC result re-read as NCHW, separate bias pass, materialised fp32
             attention scores, true-fp32 cuBLAS), cuDNN graphs built ONCE per shape, no host syncs / host round trips
  R-fp16     fp16 I/O + fp32 accumulate everywhere: channels-last cuDNN convs, fp16 cuBLAS, fused library SDPA
  R-literal  R-cached arithmetic with the reference's per-call behaviour: a new cuDNN graph built for every conv call
             (`conv2d.py:10-23`) and the device synchronisations of `ff/group_norm.py:6`, `attention/sdpa.py:70-71`,
             `variants/sd.py:40-41`; eager (cannot be captured), a bounded number of steps

The reference cannot be imported offline (CuPy / tinygrad missing), so torch tensors are the containers -
as in the reference's own tests (`tests/conv2d.py:19-22`). LayerNorm uses the library kernel with canonical strides: cuDNN
rejects the reference's stride declaration at batch > 1 (tests/golden/cudnn_layernorm_probe.json).
"""
import math

import torch
import torch.nn.functional as F

from tinyfusers_b200.synthetic import UNET_INPUT_BLOCKS, UNET_MIDDLE_BLOCK, UNET_OUTPUT_BLOCKS


class _ConvGraphs:
    """cuDNN python-frontend conv_fprop graphs, cached per (shape, dtype, padding, stride) unless literal."""

    def __init__(self, literal=False):
        import cudnn
        self.cudnn = cudnn
        self.handle = cudnn.create_handle()
        self.cache = {}
        self.literal = literal
        self.builds = 0

    def bind_stream(self):
        self.cudnn.set_stream(handle=self.handle, stream=torch.cuda.current_stream().cuda_stream)

    def _build(self, xs, ws, padding, stride, half, channels_last):
        cudnn = self.cudnn
        dt = cudnn.data_type.HALF if half else cudnn.data_type.FLOAT
        g = cudnn.pygraph(handle=self.handle, name="conv2d", io_data_type=dt, intermediate_data_type=dt,
                          compute_data_type=cudnn.data_type.FLOAT)
        N, C, H, W = xs
        K, _, R, S = ws
        if channels_last:
            x_stride, w_stride = [H * W * C, 1, W * C, C], [C * R * S, 1, S * C, C]
        else:   # the reference's explicit NCHW / KCRS strides (conv2d.py:15-16)
            x_stride, w_stride = [C * H * W, H * W, W, 1], [C * R * S, R * S, S, 1]
        X = g.tensor(name="X", dim=list(xs), stride=x_stride, data_type=dt)
        Wt = g.tensor(name="W", dim=list(ws), stride=w_stride, data_type=dt)
        Y = g.conv_fprop(image=X, weight=Wt, padding=list(padding), stride=list(stride), dilation=[1, 1],
                         compute_data_type=cudnn.data_type.FLOAT)
        Y.set_output(True)
        g.build([cudnn.heur_mode.A])
        self.builds += 1
        ydim = tuple(Y.get_dim())
        ws_t = torch.empty(max(g.get_workspace_size(), 1), dtype=torch.uint8, device="cuda")
        return g, X, Wt, Y, ydim, ws_t

    def __call__(self, x, w, padding, stride, channels_last):
        half = x.dtype == torch.float16
        key = (tuple(x.shape), tuple(w.shape), tuple(padding), tuple(stride), half, channels_last)
        ent = None if self.literal else self.cache.get(key)
        if ent is None:
            ent = self._build(key[0], key[1], padding, stride, half, channels_last)
            if not self.literal:
                self.cache[key] = ent
        g, X, Wt, Y, ydim, ws_t = ent
        N, K, Ho, Wo = ydim
        y = torch.empty((N, Ho, Wo, K), dtype=x.dtype, device=x.device)   # the frontend's default output layout is NHWC
        g.execute({X: x, Wt: w, Y: y}, ws_t, handle=self.handle)
        return y.permute(0, 3, 1, 2)          # re-read as NCHW (conv2d.py:27); a view, like CuPy's transpose


class LibraryStep:
    """One CFG denoising step (UNet at batch 2B + CFG combine + DDIM) through cuDNN / cuBLAS / library kernels."""

    def __init__(self, sd, B_img, HW, variant="cached", conv="pygraph", prefix="model.diffusion_model"):
        assert variant in ("cached", "fp16", "literal")
        self.variant, self.B, self.HW, self.P = variant, B_img, HW, prefix
        self.half = variant == "fp16"
        self.dt = torch.float16 if self.half else torch.float32
        self.cl = self.half                       # channels-last activations / weights in the fp16 variant
        self.literal = variant == "literal"
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.conv_backend = conv
        self.graphs = None
        if conv == "pygraph":
            self.graphs = _ConvGraphs(literal=self.literal)
        self.w = {}
        for k, v in sd.items():
            if not k.startswith(prefix):
                continue
            t = v.to(dev, self.dt)
            if t.dim() == 4 and self.cl:
                t = t.contiguous(memory_format=torch.channels_last)
            self.w[k] = t
        self.latent = torch.zeros((B_img, 4, HW, HW), dtype=torch.float32, device=dev)
        self.context = torch.zeros((2 * B_img, 77, 768), dtype=torch.float32, device=dev)   # [uncond ; cond]
        self.tab = torch.ones((3, 1024), dtype=torch.float32, device=dev)                    # [t ; a_t ; a_prev]
        self.idx = torch.zeros(1, dtype=torch.long, device=dev)
        self.guidance = 7.5
        i = torch.arange(160, dtype=torch.float32, device=dev)
        self.freqs = torch.exp(-math.log(10000.0) * i / 160.0)

    # ---- operators ----
    def _sync(self):
        if self.literal:
            torch.cuda.synchronize()

    def conv(self, x, p, padding=(0, 0), stride=(1, 1)):
        w, b = self.w[p + ".weight"], self.w[p + ".bias"]
        if self.graphs is not None:
            if not self.cl:
                x = x.contiguous()
            else:
                x = x.contiguous(memory_format=torch.channels_last)
            y = self.graphs(x, w, padding, stride, self.cl)
        else:
            y = F.conv2d(x, w, None, stride, padding)
        return y + b.reshape(1, -1, 1, 1)         # separate bias pass (conv2d.py:58)

    def gn(self, x, p, eps=1e-5):
        self._sync()                              # ff/group_norm.py:6
        return F.group_norm(x, 32, self.w[p + ".weight"], self.w[p + ".bias"], eps)

    def ln(self, x, p):
        return F.layer_norm(x, (x.shape[-1],), self.w[p + ".weight"], self.w[p + ".bias"], 1e-5)

    def linear(self, x, p, bias=True):
        return F.linear(x, self.w[p + ".weight"], self.w[p + ".bias"] if bias else None)

    def sdpa(self, q, k, v):
        if self.half:
            return F.scaled_dot_product_attention(q, k, v)
        s = (1.0 / math.sqrt(q.shape[-1])) * torch.matmul(q, k.transpose(-1, -2))    # sdpa.py:66, scores materialised
        self._sync()                                                                  # sdpa.py:70-71
        return torch.matmul(torch.softmax(s, dim=-1), v)

    def attn(self, x, ctx, p, nh, d):
        B, T, C = x.shape
        q = self.linear(x, p + ".to_q", False).reshape(B, T, nh, d).transpose(1, 2)
        k = self.linear(ctx, p + ".to_k", False).reshape(B, -1, nh, d).transpose(1, 2)
        v = self.linear(ctx, p + ".to_v", False).reshape(B, -1, nh, d).transpose(1, 2)
        o = self.sdpa(q, k, v).contiguous().reshape(B, T, nh * d)      # head-major reshape, attention.py:39
        return self.linear(o, p + ".to_out.0")

    def st(self, x, ctx, p, nh, d):
        B, C, H, W = x.shape
        h = self.conv(self.gn(x, p + ".norm"), p + ".proj_in")
        h = h.reshape(B, C, H * W).transpose(1, 2)
        t = p + ".transformer_blocks.0"
        h = self.attn(self.ln(h, t + ".norm1"), self.ln(h, t + ".norm1"), t + ".attn1", nh, d) + h
        h = self.attn(self.ln(h, t + ".norm2"), ctx, t + ".attn2", nh, d) + h
        g = self.linear(self.ln(h, t + ".norm3"), t + ".ff.net.0.proj")
        a, gate = g.chunk(2, dim=-1)
        h = self.linear(a * F.gelu(gate, approximate="tanh"), t + ".ff.net.2") + h
        h = h.transpose(1, 2).reshape(B, C, H, W)
        return self.conv(h, p + ".proj_out") + x

    def res(self, x, emb, p, cin, cout):
        h = self.conv(F.silu(self.gn(x, p + ".in_layers.0")), p + ".in_layers.2", (1, 1))
        h = h + self.linear(F.silu(emb), p + ".emb_layers.1").reshape(1, -1, 1, 1)
        h = self.conv(F.silu(self.gn(h, p + ".out_layers.0")), p + ".out_layers.3", (1, 1))
        return (self.conv(x, p + ".skip_connection") if cin != cout else x) + h

    def layer(self, x, emb, ctx, p, layer):
        kind = layer[0]
        if kind == "conv":
            return self.conv(x, p, (1, 1))
        if kind == "res":
            return self.res(x, emb, p, layer[1], layer[2])
        if kind == "st":
            return self.st(x, ctx, p, layer[2], layer[3])
        if kind == "down":
            return self.conv(x, p + ".op", (1, 1), (2, 2))
        if kind == "up":
            return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), p + ".conv", (1, 1))
        raise ValueError(kind)

    def unet(self, x, t, ctx):
        P = self.P
        ang = t.reshape(1, 1) * self.freqs.reshape(1, -1)
        temb = torch.cat((torch.cos(ang), torch.sin(ang)), dim=-1).to(self.dt)
        emb = self.linear(F.silu(self.linear(temb, P + ".time_embed.0")), P + ".time_embed.2")
        saved = []
        for i, block in enumerate(UNET_INPUT_BLOCKS):
            for j, layer in enumerate(block):
                x = self.layer(x, emb, ctx, f"{P}.input_blocks.{i}.{j}", layer)
            saved.append(x)
        for j, layer in enumerate(UNET_MIDDLE_BLOCK):
            x = self.layer(x, emb, ctx, f"{P}.middle_block.{j}", layer)
        for i, block in enumerate(UNET_OUTPUT_BLOCKS):
            x = torch.cat((x, saved.pop()), dim=1)
            for j, layer in enumerate(block):
                x = self.layer(x, emb, ctx, f"{P}.output_blocks.{i}.{j}", layer)
        return self.conv(F.silu(self.gn(x, P + ".out.0")), P + ".out.2", (1, 1))

    def step(self):
        """variants/sd.py:27-59 with the schedule read from device tables (graph-capturable)."""
        B = self.B
        t = self.tab[0].index_select(0, self.idx)
        a_t = self.tab[1].index_select(0, self.idx)
        a_prev = self.tab[2].index_select(0, self.idx)
        x = torch.cat((self.latent, self.latent), dim=0).to(self.dt)
        if self.cl:
            x = x.contiguous(memory_format=torch.channels_last)
        if self.literal:                          # variants/sd.py:34-41: host round trip of the batch + two syncs
            x = x.cpu().to(self.dev)
            torch.cuda.synchronize()
        out = self.unet(x, t, self.context.to(self.dt)).float()
        e_t = out[:B] + self.guidance * (out[B:] - out[:B])
        pred_x0 = (self.latent - torch.sqrt(1.0 - a_t) * e_t) / torch.sqrt(a_t)
        self.latent.copy_(torch.sqrt(a_prev) * pred_x0 + torch.sqrt(1.0 - a_prev) * e_t)

    # ---- measurement ----
    def load(self, unc, ctx, lat, ts, alphas, alphas_prev, guidance):
        self.latent.copy_(lat)
        self.context[:self.B].copy_(unc)
        self.context[self.B:].copy_(ctx)
        n = len(ts)
        self.tab[0, :n].copy_(torch.tensor([float(v) for v in ts]))
        self.tab[1, :n].copy_(alphas.float())
        self.tab[2, :n].copy_(alphas_prev.float())
        self.idx.fill_(n // 2)
        self.guidance = float(guidance)

    def time_ms(self, steps, warmup=3, capture=True):
        """-> (ms per step, 'graph' | 'eager')."""
        tf32_mm, tf32_cd, bench = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark
        torch.backends.cuda.matmul.allow_tf32 = False     # CuPy's cuBLAS default (CUPY_TF32=0): true fp32 SGEMM
        torch.backends.cudnn.allow_tf32 = True            # the reference's conv graph carries no numerical-note filter
        torch.backends.cudnn.benchmark = True
        try:
            with torch.no_grad():
                lat0 = self.latent.clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    if self.graphs is not None:
                        self.graphs.bind_stream()
                    for _ in range(warmup):
                        self.step()
                    side.synchronize()
                    mode, g = "eager", None
                    if capture and not self.literal:
                        try:
                            g = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(g, stream=side):
                                self.step()
                            mode = "graph"
                            g.replay()
                            side.synchronize()
                        except Exception:
                            g, mode = None, "eager (capture failed)"
                            torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(side)
                    for _ in range(steps):
                        if g is not None:
                            g.replay()
                        else:
                            self.step()
                    b.record(side)
                    side.synchronize()
                    ms = a.elapsed_time(b) / steps
                torch.cuda.current_stream().wait_stream(side)
                finite = bool(torch.isfinite(self.latent).all())
                self.latent.copy_(lat0)
            return ms, mode, finite
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = tf32_mm, tf32_cd, bench


def measure(sd, inputs, schedule, B_img=1, HW=64, steps=10, literal_steps=1):
    """-> dict for bench.py's `library_baseline` block. `sd`: fp32 CPU state dict; inputs = (lat, unc, ctx); schedule = (ts, a, ap)."""
    lat, unc, ctx = inputs
    ts, alphas, alphas_prev = schedule
    res = {"what": "the same CFG UNet step through cuDNN conv_fprop graphs (python frontend, the reference's own declarations) + "
                   "cuBLAS + library kernels on this GPU, CUDA-graph captured (baseline/ref_cudnn.py; BASELINE.md section 3)",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    plans = [("R_cached_fp32", "cached", "pygraph", steps), ("R_fp16", "fp16", "pygraph", steps), ("R_fp16_torch_conv", "fp16", "torch", steps)]
    if literal_steps > 0:
        plans.append(("R_literal_fp32", "literal", "pygraph", literal_steps))
    for name, variant, conv, n in plans:
        try:
            try:
                lib = LibraryStep(sd, B_img, HW, variant, conv)
            except Exception as exc:   # python frontend unusable: the same cuDNN engines through torch's binding
                if conv != "pygraph":
                    raise
                lib = LibraryStep(sd, B_img, HW, variant, "torch")
                res.setdefault("notes", []).append(f"{name}: cudnn python frontend unavailable ({type(exc).__name__}), F.conv2d used")
            lib.load(unc, ctx, lat, ts, alphas, alphas_prev, 7.5)
            try:
                ms, mode, finite = lib.time_ms(n, warmup=1 if variant == "literal" else 3)
            except Exception as exc:
                if lib.conv_backend != "pygraph":
                    raise
                res.setdefault("notes", []).append(f"{name}: pygraph path failed ({type(exc).__name__}: {str(exc)[:120]}), F.conv2d used")
                torch.cuda.synchronize()
                lib = LibraryStep(sd, B_img, HW, variant, "torch")
                lib.load(unc, ctx, lat, ts, alphas, alphas_prev, 7.5)
                ms, mode, finite = lib.time_ms(n, warmup=1 if variant == "literal" else 3)
            res[name] = {"ms_per_step": ms, "steps_per_s": 1000.0 / ms, "mode": mode, "conv": lib.conv_backend,
                         "steps_timed": n, "finite": finite,
                         "cudnn_graph_builds": lib.graphs.builds if lib.graphs is not None else None}
            del lib
            torch.cuda.empty_cache()
        except Exception as exc:
            res[name] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
            torch.cuda.synchronize()
    return res
