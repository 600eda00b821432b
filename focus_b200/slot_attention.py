"""B200-native drop-in for FOCUS's video slot-attention encoder.

Mirrors the nn.Module contract of the reference
    /root/reference/slowfast/models/STEVE/steve.py:11-105  class SlotAttentionVideo
(constructor signature, attribute names, state_dict keys and shapes, parameter
creation / RNG order, `forward(inputs[B,T,N,D]) -> (slots[B,T,K,Ds], attns[B,T,N,K])`)
so it can replace `STEVEEncoder.savi` (steve.py:229-232) unchanged.  The
computation itself is hand-written sm_100a CUDA behind a C ABI
(include/focus_savi.h); the sub-modules declared here only hold parameters.
There is no CPU / PyTorch fallback: a missing library or a non-CUDA input raises.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib


# ---- parameter containers (names/shapes/init = steve.py:28-50, utils.py:92-118, transformer.py:8-20,56-68,91-103)
def _linear(n_in, n_out, bias=True, weight_init="xavier", gain=1.0):
    m = nn.Linear(n_in, n_out, bias)
    if weight_init == "kaiming":
        nn.init.kaiming_uniform_(m.weight, nonlinearity="relu")
    else:
        nn.init.xavier_uniform_(m.weight, gain)
    if bias:
        nn.init.zeros_(m.bias)
    return m


def _gru_cell(n_in, n_hidden):
    m = nn.GRUCell(n_in, n_hidden, True)
    nn.init.xavier_uniform_(m.weight_ih)
    nn.init.orthogonal_(m.weight_hh)
    nn.init.zeros_(m.bias_ih)
    nn.init.zeros_(m.bias_hh)
    return m


class _Params(nn.Module):
    """A parameter holder: never called, the CUDA kernels read its tensors."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("focus_b200 parameter containers are not callable; use SlotAttentionVideo.forward")


class _PredictorAttention(_Params):
    def __init__(self, d_model, gain):
        super().__init__()
        self.proj_q = _linear(d_model, d_model, bias=False)
        self.proj_k = _linear(d_model, d_model, bias=False)
        self.proj_v = _linear(d_model, d_model, bias=False)
        self.proj_o = _linear(d_model, d_model, bias=False, gain=gain)


class _PredictorBlock(_Params):
    def __init__(self, d_model, gain):
        super().__init__()
        self.attn_layer_norm = nn.LayerNorm(d_model)
        self.attn = _PredictorAttention(d_model, gain)
        self.ffn_layer_norm = nn.LayerNorm(d_model)
        # indices 0 and 2 carry the weights, as in the reference's nn.Sequential(linear, ReLU, linear, Dropout)
        self.ffn = nn.Sequential(_linear(d_model, 4 * d_model, weight_init="kaiming"), nn.Identity(),
                                 _linear(4 * d_model, d_model, gain=gain))


class _Predictor(_Params):
    def __init__(self, num_blocks, d_model):
        super().__init__()
        gain = (2 * num_blocks) ** (-0.5) if num_blocks > 0 else 1.0
        self.blocks = nn.ModuleList([_PredictorBlock(d_model, gain) for _ in range(num_blocks)])
        self.layer_norm = nn.LayerNorm(d_model)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _SaviFunction(torch.autograd.Function):
    """autograd.Function over the C ABI: savi_pack_params + savi_forward / savi_backward."""

    @staticmethod
    def forward(ctx, shape, grad_sync, drop, inputs, noise, *params):
        dev = inputs.device
        sizes = _lib.query(shape)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        B, T, N, K, Ds = shape.B, shape.T, shape.N, shape.K, shape.Ds
        u8 = dict(dtype=torch.uint8, device=dev)
        packed = torch.empty(sizes.packed_bytes, **u8)
        saved = torch.empty(sizes.saved_bytes, **u8)
        ws = torch.empty(max(sizes.fwd_ws_bytes, 16), **u8)
        slots = torch.empty(B, T, K, Ds, dtype=torch.float32, device=dev)
        attn = torch.empty(B, T, N, K, dtype=inputs.dtype, device=dev)
        p32 = [p.detach() if p.dtype == torch.float32 else p.detach().float() for p in params]
        p32 = [p if p.is_contiguous() else p.contiguous() for p in p32]
        ptrs = (ctypes.c_void_p * len(p32))(*[p.data_ptr() for p in p32])
        _lib.check(_lib.lib.savi_pack_params(ctypes.byref(shape), ptrs, _ptr(packed), stream), "savi_pack_params")
        launches = _lib.lib.savi_last_launch_count()
        _lib.check(_lib.lib.savi_forward(ctypes.byref(shape), _ptr(packed), _ptr(inputs), _ptr(noise), _ptr(slots),
                                         _ptr(attn), _ptr(saved), _ptr(ws), _ptr(drop), stream), "savi_forward")
        _SaviFunction.last_launches = launches + _lib.lib.savi_last_launch_count()
        ctx.shape = shape
        ctx.grad_sync = grad_sync
        ctx.sizes = sizes
        ctx.param_meta = [(p.shape, p.dtype) for p in params]
        ctx.drop = drop
        ctx.save_for_backward(inputs, noise, packed, saved)
        ctx.set_materialize_grads(False)
        return (slots if inputs.dtype == torch.float32 else slots.to(inputs.dtype)), attn

    @staticmethod
    def backward(ctx, g_slots, g_attn):
        inputs, noise, packed, saved = ctx.saved_tensors
        shape, sizes = ctx.shape, ctx.sizes
        dev = inputs.device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if g_slots is None:
            g_slots = torch.zeros(shape.B, shape.T, shape.K, shape.Ds, dtype=torch.float32, device=dev)
        g_slots = g_slots.float().contiguous()
        if g_attn is not None:
            g_attn = g_attn.to(inputs.dtype).contiguous()
        ws = torch.empty(max(sizes.bwd_ws_bytes, 16), dtype=torch.uint8, device=dev)
        g_in = torch.empty_like(inputs)
        if ctx.grad_sync is not None and hasattr(ctx.grad_sync, "buffer"):
            g_par = ctx.grad_sync.buffer(sizes.param_floats, dev)        # symmetric memory: the peers read it directly
        else:
            g_par = torch.empty(sizes.param_floats, dtype=torch.float32, device=dev)
        g_noise = torch.empty_like(noise) if ctx.needs_input_grad[4] else None
        _lib.check(_lib.lib.savi_backward(ctypes.byref(shape), _ptr(packed), _ptr(inputs), _ptr(noise), _ptr(saved),
                                          _ptr(g_slots), _ptr(g_attn), _ptr(g_in), _ptr(g_par), _ptr(g_noise),
                                          _ptr(ws), _ptr(ctx.drop), stream), "savi_backward")
        _SaviFunction.last_launches = _lib.lib.savi_last_launch_count()
        if ctx.grad_sync is not None:                # data parallel: ONE all-reduce of the flat buffer (focus_b200/distributed.py)
            g_par = ctx.grad_sync(g_par)
        off, num = _lib.param_layout(shape, len(ctx.param_meta))
        grads = []
        for (shp, dt), o, n in zip(ctx.param_meta, off, num):
            g = g_par[o:o + n].view(shp)           # every parameter gets a tensor (zeros when unused): DDP-safe
            grads.append(g if dt == torch.float32 else g.to(dt))
        return (None, None, None, g_in if ctx.needs_input_grad[3] else None, g_noise) + tuple(grads)


_SaviFunction.last_launches = 0


class SlotAttentionVideo(nn.Module):
    """Same constructor and forward contract as the reference module (steve.py:13-18, 52)."""

    def __init__(self, num_iterations, num_slots, input_size, slot_size, mlp_hidden_size,
                 num_predictor_blocks=1, num_predictor_heads=4, dropout=0.1, epsilon=1e-8):
        super().__init__()
        self.num_iterations = num_iterations
        self.num_slots = num_slots
        self.input_size = input_size
        self.slot_size = slot_size
        self.mlp_hidden_size = mlp_hidden_size
        self.epsilon = epsilon
        self.num_predictor_blocks = num_predictor_blocks
        self.num_predictor_heads = num_predictor_heads
        self.dropout = dropout
        self.cluster = 0               # CTAs per clip; 0 lets the library choose
        self.grad_sync = None          # focus_b200.distributed.attach_grad_sync: flat gradient all-reduce inside backward
        # Cast policy (the reference follows `inputs` / autocast, steve_train_net.py:95).  None: compute in the dtype of
        # `inputs` (fp32 -> fp32 kernels, 1e-5 parity; bf16 -> tensor-core kernels, 2e-2 parity; fp16, what autocast hands
        # over, is computed as bf16).  torch.bfloat16: fp32 inputs are cast on entry and the outputs cast back, so the
        # stock fp32 trainer (every FOCUS yaml has TRAIN.MIXED_PRECISION False) gets the tcgen05 kernels; opt-in because
        # it trades the 1e-5 parity for the bf16 one.
        self.compute_dtype = None

        # creation order == reference order, so torch.manual_seed(s) gives identical initial weights
        self.slot_mu = nn.Parameter(torch.Tensor(1, 1, slot_size))
        self.slot_log_sigma = nn.Parameter(torch.Tensor(1, 1, slot_size))
        nn.init.xavier_uniform_(self.slot_mu)
        nn.init.xavier_uniform_(self.slot_log_sigma)
        self.norm_inputs = nn.LayerNorm(input_size)
        self.norm_slots = nn.LayerNorm(slot_size)
        self.norm_mlp = nn.LayerNorm(slot_size)
        self.project_q = _linear(slot_size, slot_size, bias=False)
        self.project_k = _linear(input_size, slot_size, bias=False)
        self.project_v = _linear(input_size, slot_size, bias=False)
        self.gru = _gru_cell(slot_size, slot_size)
        self.mlp = nn.Sequential(_linear(slot_size, mlp_hidden_size, weight_init="kaiming"), nn.Identity(),
                                 _linear(mlp_hidden_size, slot_size))
        if num_predictor_blocks > 0:
            assert slot_size % num_predictor_heads == 0, "d_model must be divisible by num_heads"
        self.predictor = _Predictor(num_predictor_blocks, slot_size)

    def _dropout_active(self):
        return self.training and self.dropout > 0 and self.num_predictor_blocks > 0

    def _draw_dropout_masks(self, B, T, like, sizes):
        """Training-mode predictor dropout (reference transformer.py:12-13, 44, 48, 68; constructor default 0.1, steve.py:16).
        The kernels multiply by masks drawn HERE with the reference's own RNG consumption: the reference runs the predictor
        after EVERY frame (the result after the last one is discarded, steve.py:100), and each block draws, in this order,
        F.dropout on the attention probabilities [B,H,K,K], on the proj_o output [B,K,Ds] and on the FFN output [B,K,Ds].
        Calling torch's own dropout on ones of those shapes and dtypes consumes the generator exactly as the reference does
        (same kernel, same launch geometry), so with equal seeds the masks ARE the reference's
        (tests/test_reference_integration.py).  Layout: include/focus_savi.h, savi_forward."""
        import torch.nn.functional as Fn
        p, H, K, Ds, nb = self.dropout, self.num_predictor_heads, self.num_slots, self.slot_size, self.num_predictor_blocks
        amp = torch.is_autocast_enabled("cuda")
        att_dt = torch.float32 if amp else like.dtype            # softmax output (autocast: an fp32 op)
        lin_dt = torch.get_autocast_dtype("cuda") if amp else like.dtype     # nn.Linear outputs
        Sp = (T - 1) * nb
        dev = like.device
        att = torch.empty(max(Sp, 1), B, H, K, K, dtype=torch.float32, device=dev)
        out = torch.empty(max(Sp, 1), B, K, Ds, dtype=torch.float32, device=dev)
        ffn = torch.empty(max(Sp, 1), B, K, Ds, dtype=torch.float32, device=dev)
        one_att = torch.ones(B, H, K, K, dtype=att_dt, device=dev)
        one_lin = torch.ones(B, K, Ds, dtype=lin_dt, device=dev)
        for t in range(T):
            for j in range(nb):
                m_a, m_o, m_f = Fn.dropout(one_att, p, True), Fn.dropout(one_lin, p, True), Fn.dropout(one_lin, p, True)
                if t < T - 1:                                   # the draw after the last frame only advances the generator
                    f = j * (T - 1) + t
                    att[f], out[f], ffn[f] = m_a, m_o, m_f
        if Sp == 0:
            return None
        flat = torch.cat([att.reshape(-1), out.reshape(-1), ffn.reshape(-1)])
        assert flat.numel() == sizes.dropout_floats
        return flat

    def _ordered_params(self):
        """Parameters in the order of the flat buffer (include/focus_savi.h: savi_param_layout)."""
        p = [self.slot_mu, self.slot_log_sigma, self.norm_inputs.weight, self.norm_inputs.bias,
             self.norm_slots.weight, self.norm_slots.bias, self.norm_mlp.weight, self.norm_mlp.bias,
             self.project_q.weight, self.project_k.weight, self.project_v.weight,
             self.gru.weight_ih, self.gru.weight_hh, self.gru.bias_ih, self.gru.bias_hh,
             self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight, self.mlp[2].bias]
        for blk in self.predictor.blocks:
            p += [blk.attn_layer_norm.weight, blk.attn_layer_norm.bias,
                  blk.attn.proj_q.weight, blk.attn.proj_k.weight, blk.attn.proj_v.weight, blk.attn.proj_o.weight,
                  blk.ffn_layer_norm.weight, blk.ffn_layer_norm.bias,
                  blk.ffn[0].weight, blk.ffn[0].bias, blk.ffn[2].weight, blk.ffn[2].bias]
        p += [self.predictor.layer_norm.weight, self.predictor.layer_norm.bias]
        return p

    def make_shape(self, B, T, N, dtype):
        return _lib.SaviShape(B=B, T=T, N=N, D=self.input_size, Ds=self.slot_size, M=self.mlp_hidden_size,
                              K=self.num_slots, I=self.num_iterations, blocks=self.num_predictor_blocks,
                              heads=self.num_predictor_heads,
                              dtype=_lib.SAVI_DTYPE_F32 if dtype == torch.float32 else _lib.SAVI_DTYPE_BF16,
                              cluster=self.cluster, eps=self.epsilon, ln_eps=1e-5)

    def forward(self, inputs, noise=None, dropout_masks=None):
        """inputs [B,T,N,D] (fp32 or bf16, CUDA).  `noise` [B,K,Ds] optionally injects the
        N(0,1) slot draw (tests); by default it is drawn exactly as the reference does (steve.py:56).
        `dropout_masks` (tests) injects the flat predictor dropout-mask buffer of include/focus_savi.h; by default the masks
        are drawn with the reference's RNG consumption when the module is in training mode with dropout > 0."""
        if inputs.dim() != 4 or inputs.shape[-1] != self.input_size:
            raise ValueError("inputs must be [B, T, num_inputs, %d], got %s" % (self.input_size, tuple(inputs.shape)))
        if not inputs.is_cuda:
            raise RuntimeError("focus_b200.SlotAttentionVideo has no CPU path: inputs must live on a B200 (sm_100) device")
        if inputs.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise TypeError("inputs must be float32, bfloat16 or float16, got %s" % inputs.dtype)
        out_dtype = inputs.dtype
        cd = self.compute_dtype or (torch.bfloat16 if inputs.dtype == torch.float16 else inputs.dtype)
        if cd not in (torch.float32, torch.bfloat16):
            raise TypeError("compute_dtype must be None, torch.float32 or torch.bfloat16, got %s" % cd)
        if inputs.dtype != cd:
            inputs = inputs.to(cd)                 # differentiable: the gradient comes back in the caller's dtype
        B, T, N, _ = inputs.shape
        if noise is None:
            noise = inputs.new_empty(B, self.num_slots, self.slot_size).normal_()
        noise = noise.to(device=inputs.device, dtype=torch.float32).contiguous()
        inputs = inputs.contiguous()
        shape = self.make_shape(B, T, N, inputs.dtype)
        drop = None
        if dropout_masks is not None:
            drop = dropout_masks.to(device=inputs.device, dtype=torch.float32).contiguous()
            if drop.numel() != _lib.query(shape).dropout_floats:
                raise ValueError("dropout_masks must hold %d floats" % _lib.query(shape).dropout_floats)
        elif self._dropout_active():                  # after the slot-noise draw, as in the reference (steve.py:56 precedes :100)
            with torch.cuda.device(inputs.device):
                drop = self._draw_dropout_masks(B, T, inputs, _lib.query(shape))
        with torch.cuda.device(inputs.device), torch.autocast("cuda", enabled=False):
            slots, attns = _SaviFunction.apply(shape, self.grad_sync if torch.is_grad_enabled() else None, drop, inputs, noise,
                                               *self._ordered_params())
        if out_dtype != cd:
            slots, attns = slots.to(out_dtype), attns.to(out_dtype)
        if torch.is_autocast_enabled("cuda"):
            # Output dtypes of the reference under torch.autocast("cuda") (tools/steve_train_net.py:95; probed on a B200,
            # tests/test_reference_integration.py): `slots` leave nn.GRUCell / nn.Linear, which autocast runs in the
            # autocast dtype, `attns` leave F.softmax, which autocast runs in float32 — whatever dtype `inputs` had.
            slots, attns = slots.to(torch.get_autocast_dtype("cuda")), attns.float()
        return slots, attns
