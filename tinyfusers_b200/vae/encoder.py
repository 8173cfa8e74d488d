"""CLIP text encoder + VAE image encoder containers with the reference's signatures
(reference: tinyfusers/vae/encoder.py:12-81).

CLIPTextTransformer: token + position embedding (one gather kernel) -> 12 pre-LN layers
[LayerNorm -> causal 12-head attention -> +residual ; LayerNorm -> fc1 -> quick_gelu -> fc2 -> +residual] -> final
LayerNorm. 7 launches per layer on the tcgen05 GEMM / attention kernels; the residual stream is fp16 (Tp = 80 rows,
rows >= 77 stay zero so that the swapped-operand V^T projection yields exact zeros in its pad columns).
The VAE `Encoder` (image -> latent) is not on the text-to-image path: it keeps its constructor so that checkpoint
keys load, and raises when called."""
import numpy as np
import torch

from .. import fp32, packing
from ..attention.attention import CLIPAttention
from ..ff.embedding import Embedding, _ids_tensor
from ..ff.group_norm import GroupNorm
from ..ff.layer_norm import LayerNorm
from ..ff.nn import CLIPMLP
from ..native.b200.ops import b200
from ..runtime import F16, F32, Context, as_f16, as_f32, standalone_context, stream_ptr
from ..vision.conv2d import Conv2d
from ..vision.resnet import ResnetBlock
from .mid import Mid


class Encoder:
    def __init__(self):
        sz = [(128, 128), (128, 256), (256, 512), (512, 512)]
        self.conv_in = Conv2d(3, 128, kernel_size=[3, 3], padding=[1, 1])
        arr = []
        for i, s in enumerate(sz):
            arr.append({"block": [ResnetBlock(s[0], s[1]), ResnetBlock(s[1], s[1])]})
            if i != 3:
                arr[-1]['downsample'] = {"conv": Conv2d(s[1], s[1], kernel_size=[3, 3], stride=[2, 2], padding=[0, 1, 0, 1])}
        self.down = arr
        self.mid = Mid(512)
        self.norm_out = GroupNorm(32, 512)
        self.conv_out = Conv2d(512, 8, kernel_size=[3, 3], padding=[1, 1])

    def __call__(self, x):
        raise RuntimeError("tinyfusers_b200: the VAE image Encoder is not built (text-to-image uses the Decoder only; "
                           "its asymmetric-padding stride-2 convolutions have no B200 kernel yet)")


class CLIPEncoderLayer:
    def __init__(self):
        self.self_attn = CLIPAttention()
        self.layer_norm1 = LayerNorm(768)
        self.mlp = CLIPMLP()
        self.layer_norm2 = LayerNorm(768)

    def __call__(self, hidden_states, causal_attention_mask=None):
        B, T, E = hidden_states.shape
        if fp32.enabled():
            return fp32.clip_encoder_layer(self, hidden_states)
        from ..attention.attention import _check_causal_mask
        _check_causal_mask(causal_attention_mask, T)
        ctx = standalone_context()
        Tp = (T + 7) // 8 * 8
        out = torch.empty((B, T, E), dtype=F32, device=hidden_states.device)
        for i in range(B):
            ctx.arena.reset()
            h = as_f16(hidden_states[i], rows_pad_to=8)            # (Tp, E) residual stream, pad rows zero
            xn = torch.zeros((Tp, E), dtype=F16, device=hidden_states.device)
            self._run(ctx, h.data_ptr(), xn.data_ptr(), T, Tp)
            out[i].copy_(as_f32(h)[:T])
        return out

    # h (Tp, 768) fp16 residual stream, updated in place; xn: scratch for the normalised rows (rows >= T stay zero)
    def _run(self, ctx, h_ptr, xn_ptr, T, Tp):
        self.layer_norm1._run(ctx, h_ptr, xn_ptr, 1, T, 768)
        self.self_attn._run(ctx, xn_ptr, h_ptr, T, Tp)
        self.layer_norm2._run(ctx, h_ptr, xn_ptr, 1, T, 768)
        self.mlp._run(ctx, xn_ptr, h_ptr, T)


class CLIPEncoder:
    def __init__(self):
        self.layers = [CLIPEncoderLayer() for i in range(12)]

    def __call__(self, hidden_states, causal_attention_mask=None):
        for l in self.layers:
            hidden_states = l(hidden_states, causal_attention_mask)
        return hidden_states


class CLIPTextEmbeddings:
    def __init__(self):
        self.token_embedding = Embedding(49408, 768)
        self.position_embedding = Embedding(77, 768)

    def __call__(self, input_ids, position_ids):
        return self.token_embedding(input_ids) + self.position_embedding(position_ids)


class CLIPTextTransformer:
    def __init__(self):
        self.embeddings = CLIPTextEmbeddings()
        self.encoder = CLIPEncoder()
        self.final_layer_norm = LayerNorm(768)

    def __call__(self, input_ids):
        """(B, T <= 77) token ids -> (B, T, 768) fp32 prompt embeddings (reference: encoder.py:78-81)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        b200.init(dev.index)
        ids = _ids_tensor(input_ids, dev, self.embeddings.token_embedding.weight.shape[0])
        if ids.dim() == 1:
            ids = ids.reshape(1, -1)
        B, T = ids.shape
        if T > 77:
            raise RuntimeError(f"CLIPTextTransformer: {T} tokens, the position table has 77")
        if fp32.enabled():
            return fp32.clip_text_transformer(self, ids)
        eng = self._engine(T, dev)
        out = torch.empty((B, T, 768), dtype=F32, device=dev)
        for i in range(B):
            eng["ids"].copy_(ids[i])
            eng["graph"].replay()
            out[i].copy_(eng["out"])
        return out

    def _enqueue(self, eng):
        """One prompt on the current stream: ids -> token + position rows -> 12 layers -> final LayerNorm -> fp32 (T, 768)."""
        T, Tp, h, xn = eng["T"], eng["Tp"], eng["h"], eng["xn"]
        ctx = eng["ctx"]
        tok = self.embeddings.token_embedding.weight
        pos = self.embeddings.position_embedding.weight
        S = stream_ptr()
        ctx.arena.reset()
        if not ctx.dry:
            h.zero_()                                     # pad rows (T..Tp) must be zero
            xn.zero_()
            b200.check(b200.tf_embedding_f16(eng["ids"].data_ptr(), tok.data_ptr(), pos.data_ptr(), h.data_ptr(), T, T, 768,
                                             tok.shape[0], S), "tf_embedding_f16")
        with packing.domain("clip"):      # packing the text encoder's weights must not drop the UNet sampler's graphs
            for l in self.encoder.layers:
                l._run(ctx, h.data_ptr(), xn.data_ptr(), T, Tp)
            self.final_layer_norm._run(ctx, h.data_ptr(), xn.data_ptr(), 1, T, 768)
        if not ctx.dry:
            b200.check(b200.tf_cast_f16_to_f32(xn.data_ptr(), eng["out"].data_ptr(), T * 768, S), "tf_cast_f16_to_f32")

    def _engine(self, T, dev):
        """Static buffers + the captured launch sequence (97 launches, launch-bound when issued eagerly: 1.4 ms per prompt) for
        prompts of T tokens; re-captured when the encoder's weights were repacked or replaced (packing.generation)."""
        from .. import get_layernorm_strided, get_quirks
        key = (dev.index, T, get_quirks(), get_layernorm_strided())
        engines = self.__dict__.setdefault("_engines", {})
        eng = engines.get(key)
        if eng is None:
            Tp = (T + 7) // 8 * 8
            # its own context / arena: the graph bakes in arena addresses, which must not move when some other stand-alone
            # call grows the shared scratch arena
            ectx = Context(dev, get_quirks(), get_layernorm_strided())
            ectx.ensure_workspaces()
            eng = engines[key] = {"T": T, "Tp": Tp, "ctx": ectx, "graph": None, "gen": None,
                                  "ids": torch.zeros((T,), dtype=torch.int32, device=dev),
                                  "h": torch.zeros((Tp, 768), dtype=F16, device=dev),
                                  "xn": torch.zeros((Tp, 768), dtype=F16, device=dev),
                                  "out": torch.zeros((T, 768), dtype=F32, device=dev)}
        if eng["graph"] is None or eng["gen"] != packing.generation("clip"):
            if "sized" not in eng:                        # measure the arena with a dry run, then allocate it once
                ectx = eng["ctx"]
                ectx.dry = ectx.arena.dry = True
                self._enqueue(eng)
                ectx.dry = ectx.arena.dry = False
                ectx.arena.reserve(ectx.arena.peak, dev)
                eng["sized"] = True
            self._enqueue(eng)                            # eager: weight packing never happens inside a capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(eng)
            eng["graph"], eng["gen"] = g, packing.generation("clip")
        return eng
