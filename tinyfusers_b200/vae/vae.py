"""AutoencoderKL with the reference's signature (reference: tinyfusers/vae/vae.py:5-18).

Text-to-image only needs `post_quant_conv` + `decoder` (variants/sd.py:48-54); the image encoder keeps its
constructor (checkpoint keys load) but has no B200 kernels yet and raises when called."""
import torch

from ..native.b200.ops import b200
from ..runtime import F32, require_cuda, stream_ptr
from ..vision.conv2d import Conv2d
from .decoder import Decoder
from .encoder import Encoder


class AutoencoderKL:
    def __init__(self):
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.quant_conv = Conv2d(8, 8, kernel_size=[1, 1])
        self.post_quant_conv = Conv2d(4, 4, kernel_size=[1, 1])

    def __call__(self, x):
        latent = self.encoder(x)     # raises: image encoder not built
        return self.decoder(self.post_quant(latent[:, 0:4]))

    def post_quant(self, x, scale=1.0):
        """post_quant_conv(scale * x) on the fp32 NCHW latent: 16 MACs per pixel, one launch (vae.py:10, sd.py:49)."""
        require_cuda(x, "x")
        x = x.to(F32).contiguous()
        N, C, H, W = x.shape
        w = self.post_quant_conv.weight.to(F32).reshape(self.post_quant_conv.weight.shape[0], -1).contiguous()
        b = self.post_quant_conv.bias
        out = torch.empty((N, w.shape[0], H, W), dtype=F32, device=x.device)
        st = b200.tf_conv1x1_small_f32nchw(x.data_ptr(), w.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                                           N, C, w.shape[0], H * W, float(scale), stream_ptr())
        b200.check(st, "tf_conv1x1_small_f32nchw")
        return out
