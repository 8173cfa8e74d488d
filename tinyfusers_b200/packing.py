"""One-time weight re-layout (load time, not on the step path): the reference's fp32 OIHW / (out,in)
tensors -> the fp16 layouts the kernels read through TMA. Pure data movement done with torch as the
container library."""
import torch

F16, F32 = torch.float16, torch.float32


import weakref

# Bumped whenever any packed weight is (re)built and by storage.state.update_state: captured CUDA graphs bake in the
# addresses of packed weights, so whoever holds a graph (variants.sd.SamplerEngine, vae.decoder.DecoderEngine) records the
# generation at capture time and re-captures when it has moved.
_generation = 0
# Engines that hold graphs enqueue their kernels inside `with packing.domain(name)`: a weight (re)packed there only moves that
# domain's counter, so packing the VAE decoder's weights (first decode) does not throw away the UNet sampler's captured step.
# Packing outside any domain (stand-alone module calls) and update_state move the global counter, which every holder sees.
_domain = None
_domain_gen = {}


def generation(name=None):
    """name=None: the global counter; otherwise (global, domain) - what a graph holder of that domain compares."""
    return _generation if name is None else (_generation, _domain_gen.get(name, 0))


def bump_generation():
    global _generation
    _generation += 1
    return _generation


import contextlib


@contextlib.contextmanager
def domain(name):
    global _domain
    prev, _domain = _domain, name
    try:
        yield
    finally:
        _domain = prev


def _version(t):
    try:
        return t._version
    except Exception:   # tensors created under torch.inference_mode() have no version counter
        return 0


def _same(entry_refs, tensors):
    """Identity + in-place-version comparison. Source tensors are held by weak reference, so a freed-and-reallocated
    address (same data_ptr, version 0, same shape) can never be mistaken for the tensor that was packed."""
    if len(entry_refs) != len(tensors):
        return False
    for (ref, ver), t in zip(entry_refs, tensors):
        if t is None:
            if ref is not None:
                return False
            continue
        if ref is None or ref() is not t or ver != _version(t):
            return False
    return True


def cached(obj, name, tensors, builder):
    """Cache `builder()` on `obj` until any of `tensors` is replaced or modified in place."""
    slot = "_pk_" + name
    hit = obj.__dict__.get(slot)
    if hit is not None and _same(hit[0], tensors):
        return hit[1]
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        raise RuntimeError("tinyfusers_b200: weights changed while a CUDA graph was being captured; run one eager step "
                           "(or call update_state before the first call) so that packing happens outside the capture")
    val = builder()
    if torch.cuda.is_available():
        # the packed tensors are handed to kernels as STATIC weights (TF_GEMM_W_STATIC: fetched before the dependency
        # wait of programmatic dependent launch), so they must be complete in memory before anything else is enqueued
        torch.cuda.current_stream().synchronize()
    refs = tuple((None, 0) if t is None else (weakref.ref(t), _version(t)) for t in tensors)
    obj.__dict__[slot] = (refs, val)
    if _domain is None:
        bump_generation()
    else:
        _domain_gen[_domain] = _domain_gen.get(_domain, 0) + 1
    return val


def f32(t):
    return None if t is None else t.detach().to(F32).contiguous()


def linear_weight(w, k_pad_to=8):
    w = w.detach()
    out_f, in_f = w.shape
    kp = (in_f + k_pad_to - 1) // k_pad_to * k_pad_to
    if kp == in_f:
        return w.to(F16).contiguous()
    p = torch.zeros((out_f, kp), dtype=F16, device=w.device)
    p[:, :in_f] = w
    return p


def pad_rows(w, mult):
    n = w.shape[0]
    np_ = (n + mult - 1) // mult * mult
    if np_ == n:
        return w
    p = torch.zeros((np_,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
    p[:n] = w
    return p


def conv3x3_weight(w, cin_pad_to=64, cout_pad_to=8):
    """OIHW fp32 -> (Cout_pad, 3, 3, Cin_pad) fp16 (OHWI), zero padded."""
    w = w.detach()
    O, I, kh, kw = w.shape
    Ip = (I + cin_pad_to - 1) // cin_pad_to * cin_pad_to
    Op = (O + cout_pad_to - 1) // cout_pad_to * cout_pad_to
    p = torch.zeros((Op, kh, kw, Ip), dtype=F16, device=w.device)
    p[:O, :, :, :I] = w.permute(0, 2, 3, 1)
    return p.contiguous()


def conv_up2x_weight(w, cin_pad_to=64, cout_pad_to=8):
    """OIHW fp32 3x3 weight -> (4, Cout_pad, 4 * Cin_pad) fp16 for tf_conv2d_up2x_nhwc_f16: the convolution of a nearest-neighbour
    2x upsampled image as four 2 x 2 convolutions of the original one. Output phase (a, b) = (row, column) parity; its tap (i, j)
    reads input pixel (y + a - 1 + i, x + b - 1 + j) and carries the SUM of the 3x3 taps that land on that pixel - rows
    {0} / {1, 2} for a = 0, {0, 1} / {2} for a = 1 (same for columns). Summed in fp32, rounded to fp16 once."""
    w = w.detach().to(F32)
    O, I, kh, kw = w.shape
    assert kh == 3 and kw == 3
    Ip = (I + cin_pad_to - 1) // cin_pad_to * cin_pad_to
    Op = (O + cout_pad_to - 1) // cout_pad_to * cout_pad_to
    sets = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    p = torch.zeros((4, Op, 4, Ip), dtype=F32, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for i in (0, 1):
                for j in (0, 1):
                    acc = torch.zeros((O, I), dtype=F32, device=w.device)
                    for ky in sets[a][i]:
                        for kx in sets[b][j]:
                            acc += w[:, :, ky, kx]
                    p[2 * a + b, :O, 2 * i + j, :I] = acc
    return p.reshape(4, Op, 4 * Ip).to(F16).contiguous()


def conv1x1_weight(w, cout_pad_to=8):
    w = w.detach()
    O, I = w.shape[0], w.shape[1]
    return pad_rows(linear_weight(w.reshape(O, I)), cout_pad_to)


def head_pad(w, n_heads, d, dp):
    """(n_heads*d, in) -> (n_heads*dp, in): each head's rows zero-padded to dp (attention K-dim multiple of 16)."""
    w = w.detach()
    if dp == d:
        return w.to(F16).contiguous()
    inf = w.shape[1]
    p = torch.zeros((n_heads, dp, inf), dtype=F16, device=w.device)
    p[:, :d] = w.reshape(n_heads, d, inf)
    return p.reshape(n_heads * dp, inf).contiguous()


def geglu_pack(w, b):
    """GEGLU proj (2*dout, in): rows [0,dout) = value, [dout,2dout) = gate (ff/nn.py:11 split) ->
    blocks of 32 rows = 16 value rows followed by their 16 gate rows (TF_EPI_GEGLU layout)."""
    w = w.detach()
    dout = w.shape[0] // 2
    assert dout % 16 == 0
    idx = torch.arange(dout, device=w.device).reshape(-1, 16)
    perm = torch.cat((idx, idx + dout), dim=1).reshape(-1)
    wp = w[perm].to(F16).contiguous()
    bp = None if b is None else b.detach()[perm].to(F32).contiguous()
    return wp, bp


def ln_fold(w_packed, gamma, beta, bias=None):
    """LayerNorm folded into a Linear that consumes it (tf_gemm_ex_f16): W' = W diag(gamma) (fp16), c1[n] = sum_k W'[n,k]
    (from the ROUNDED W', fp32), c2 = W beta (+ bias). `w_packed` is the already packed (rows padded / permuted) fp16 weight."""
    wf = w_packed.detach().to(F32)
    g = gamma.detach().to(F32).to(wf.device)
    b = beta.detach().to(F32).to(wf.device)
    wp = (wf * g[None, :]).to(F16).contiguous()
    c1 = wp.to(F32).sum(dim=1).contiguous()
    c2 = wf @ b
    if bias is not None:
        c2 = c2 + bias.detach().to(F32)
    return wp, c1, c2.contiguous()
