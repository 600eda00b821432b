"""numpy restatement of FOCUS's video slot-attention encoder (forward + closed-form backward).

TEST INFRASTRUCTURE ONLY.  Nothing under focus_b200/ imports this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.  The
product path is the CUDA library and fails loudly when it is missing.

What is restated (reference = /root/reference, read-only):
  * forward  : slowfast/models/STEVE/steve.py:52-105  (SlotAttentionVideo.forward)
  * predictor: slowfast/models/STEVE/transformer.py:22-49 (MHA), :70-86 (block,
               incl. the is_first "residual on the normalised input" rule), :106-114
  * GRUCell / LayerNorm semantics: torch.nn (gate order r,z,n; biased variance)
  * parameter names / shapes: steve.py:13-50, utils.py:92-118

Parity pinning: the reference ships NO tests and NO golden vectors for this
path, so this oracle is pinned against fixtures produced by
importing the unmodified reference module in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz; checked by
tests/test_oracle_golden.py).

Two algebraically identical forward forms are provided:
  folded=False  k = LN(x) Wk^T * Ds^-1/2, v = LN(x) Wv^T materialised, exactly
                the reference's operation order;
  folded=True   the order the CUDA kernels use: the K slot rows carry the
                projections (qk = Ds^-1/2 * q Wk, U = (A^T xhat / S) Wv^T), so
                only the LayerNorm'd tokens xhat are streamed per iteration.
The backward below is the closed form the CUDA backward implements (SURVEY.md
Appendix A.2 transposed to the folded form); it is validated against torch
autograd of the reference module in tests/test_oracle_golden.py.
"""
import numpy as np

LN_EPS = 1e-5


# ----------------------------------------------------------------------------
# parameter inventory (names are the reference state_dict keys)
# ----------------------------------------------------------------------------
def param_shapes(K, D, Ds, M, blocks):
    """Ordered {name: shape}; order == reference state_dict order (steve.py:28-50)."""
    s = {}
    s["slot_mu"] = (1, 1, Ds)
    s["slot_log_sigma"] = (1, 1, Ds)
    for n, d in (("norm_inputs", D), ("norm_slots", Ds), ("norm_mlp", Ds)):
        s[n + ".weight"] = (d,)
        s[n + ".bias"] = (d,)
    s["project_q.weight"] = (Ds, Ds)
    s["project_k.weight"] = (Ds, D)
    s["project_v.weight"] = (Ds, D)
    s["gru.weight_ih"] = (3 * Ds, Ds)
    s["gru.weight_hh"] = (3 * Ds, Ds)
    s["gru.bias_ih"] = (3 * Ds,)
    s["gru.bias_hh"] = (3 * Ds,)
    s["mlp.0.weight"] = (M, Ds)
    s["mlp.0.bias"] = (M,)
    s["mlp.2.weight"] = (Ds, M)
    s["mlp.2.bias"] = (Ds,)
    for j in range(blocks):
        p = "predictor.blocks.%d." % j
        s[p + "attn_layer_norm.weight"] = (Ds,)
        s[p + "attn_layer_norm.bias"] = (Ds,)
        for w in ("proj_q", "proj_k", "proj_v", "proj_o"):
            s[p + "attn." + w + ".weight"] = (Ds, Ds)
        s[p + "ffn_layer_norm.weight"] = (Ds,)
        s[p + "ffn_layer_norm.bias"] = (Ds,)
        s[p + "ffn.0.weight"] = (4 * Ds, Ds)
        s[p + "ffn.0.bias"] = (4 * Ds,)
        s[p + "ffn.2.weight"] = (Ds, 4 * Ds)
        s[p + "ffn.2.bias"] = (Ds,)
    s["predictor.layer_norm.weight"] = (Ds,)
    s["predictor.layer_norm.bias"] = (Ds,)
    return s


def random_params(K, D, Ds, M, blocks, seed=0, dtype=np.float64):
    """Non-trivial random parameters (all biases / LN affines non-degenerate)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in param_shapes(K, D, Ds, M, blocks).items():
        if name.endswith("norm.weight") or name in ("norm_inputs.weight", "norm_slots.weight", "norm_mlp.weight"):
            a = 1.0 + 0.2 * rng.standard_normal(shp)
        elif len(shp) == 1 or name.startswith("slot_"):
            a = 0.2 * rng.standard_normal(shp)
        else:
            a = rng.standard_normal(shp) / np.sqrt(shp[-1])
        out[name] = a.astype(dtype)
    return out


# ----------------------------------------------------------------------------
# small ops
# ----------------------------------------------------------------------------
def _ln(x, g, b):
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(LN_EPS))
    z = xc * rstd
    return z * g + b, z, rstd


def _ln_bwd(dy, z, rstd, g):
    dz = dy * g
    dx = rstd * (dz - dz.mean(-1, keepdims=True) - z * (dz * z).mean(-1, keepdims=True))
    red = tuple(range(dy.ndim - 1))
    return dx, (dy * z).sum(red), dy.sum(red)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _softmax(x):
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(-1, keepdims=True)


def round_bf16(a):
    """Round-to-nearest-even to bfloat16 precision (returned in a's dtype)."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).astype(a.dtype).reshape(a.shape)


def _acc(d, k, v):
    d[k] = d.get(k, 0) + v


# ----------------------------------------------------------------------------
# predictor (transformer.py:22-49, 70-86, 106-114).  Dropout is the identity in evaluation; in training mode it is three
# multiplicative masks per block (attention probabilities :44, proj_o output :48, FFN output :68), which the caller passes
# in (`drop` = (att [B,H,K,K], out [B,K,Ds], ffn [B,K,Ds]) per block, entries 0 or 1/(1-p)): the oracle has no RNG.
# ----------------------------------------------------------------------------
def _predictor_fwd(P, x, blocks, heads, drop=None):
    B, K, Ds = x.shape
    dh = Ds // heads
    scale = x.dtype.type(dh ** -0.5)
    sv = []
    for j in range(blocks):
        p = "predictor.blocks.%d." % j
        y, z1, r1 = _ln(x, P[p + "attn_layer_norm.weight"], P[p + "attn_layer_norm.bias"])
        Q = (y @ P[p + "attn.proj_q.weight"].T).reshape(B, K, heads, dh).transpose(0, 2, 1, 3) * scale
        Kk = (y @ P[p + "attn.proj_k.weight"].T).reshape(B, K, heads, dh).transpose(0, 2, 1, 3)
        V = (y @ P[p + "attn.proj_v.weight"].T).reshape(B, K, heads, dh).transpose(0, 2, 1, 3)
        att = _softmax(Q @ Kk.transpose(0, 1, 3, 2))                       # [B,H,K,K]
        m_att, m_out, m_ffn = drop[j] if drop is not None else (1.0, 1.0, 1.0)
        O = ((att * m_att) @ V).transpose(0, 2, 1, 3).reshape(B, K, Ds)
        mha = (O @ P[p + "attn.proj_o.weight"].T) * m_out
        x1 = (y if j == 0 else x) + mha                                     # transformer.py:75-82
        l2, z2, r2 = _ln(x1, P[p + "ffn_layer_norm.weight"], P[p + "ffn_layer_norm.bias"])
        f = np.maximum(l2 @ P[p + "ffn.0.weight"].T + P[p + "ffn.0.bias"], 0)
        x2 = x1 + (f @ P[p + "ffn.2.weight"].T + P[p + "ffn.2.bias"]) * m_ffn
        sv.append(dict(y=y, z1=z1, r1=r1, Q=Q, Kk=Kk, V=V, att=att, O=O, l2=l2, z2=z2, r2=r2, f=f, drop=(m_att, m_out, m_ffn)))
        x = x2
    out, zf, rf = _ln(x, P["predictor.layer_norm.weight"], P["predictor.layer_norm.bias"])
    return out, dict(blocks=sv, zf=zf, rf=rf)


def _predictor_bwd(P, sv, dout, blocks, heads, G):
    B, K, Ds = dout.shape
    dh = Ds // heads
    scale = dout.dtype.type(dh ** -0.5)
    dx, dg, db = _ln_bwd(dout, sv["zf"], sv["rf"], P["predictor.layer_norm.weight"])
    _acc(G, "predictor.layer_norm.weight", dg)
    _acc(G, "predictor.layer_norm.bias", db)
    for j in reversed(range(blocks)):
        p = "predictor.blocks.%d." % j
        s = sv["blocks"][j]
        m_att, m_out, m_ffn = s["drop"]
        dx2 = dx
        dff = dx2 * m_ffn                                                  # gradient of the (dropped) FFN output
        _acc(G, p + "ffn.2.bias", dff.sum((0, 1)))
        _acc(G, p + "ffn.2.weight", np.einsum("bko,bkc->oc", dff, s["f"]))
        df = (dff @ P[p + "ffn.2.weight"]) * (s["f"] > 0)
        _acc(G, p + "ffn.0.bias", df.sum((0, 1)))
        _acc(G, p + "ffn.0.weight", np.einsum("bko,bkc->oc", df, s["l2"]))
        dl2 = df @ P[p + "ffn.0.weight"]
        d, dg, db = _ln_bwd(dl2, s["z2"], s["r2"], P[p + "ffn_layer_norm.weight"])
        _acc(G, p + "ffn_layer_norm.weight", dg)
        _acc(G, p + "ffn_layer_norm.bias", db)
        dx1 = dx2 + d
        # MHA backward
        dmo = dx1 * m_out                                                  # gradient of the (dropped) proj_o output
        _acc(G, p + "attn.proj_o.weight", np.einsum("bko,bkc->oc", dmo, s["O"]))
        dO = (dmo @ P[p + "attn.proj_o.weight"]).reshape(B, K, heads, dh).transpose(0, 2, 1, 3)
        datt = (dO @ s["V"].transpose(0, 1, 3, 2)) * m_att
        dV = (s["att"] * m_att).transpose(0, 1, 3, 2) @ dO
        dlog = s["att"] * (datt - (s["att"] * datt).sum(-1, keepdims=True))
        dQ = (dlog @ s["Kk"]) * scale
        dKk = dlog.transpose(0, 1, 3, 2) @ s["Q"]                            # Q already scaled
        flat = lambda a: a.transpose(0, 2, 1, 3).reshape(B, K, Ds)
        dQ, dKk, dV = flat(dQ), flat(dKk), flat(dV)
        dy = 0
        for nm, dd in (("proj_q", dQ), ("proj_k", dKk), ("proj_v", dV)):
            _acc(G, p + "attn." + nm + ".weight", np.einsum("bko,bkc->oc", dd, s["y"]))
            dy = dy + dd @ P[p + "attn." + nm + ".weight"]
        if j == 0:
            dy = dy + dx1
        d, dg, db = _ln_bwd(dy, s["z1"], s["r1"], P[p + "attn_layer_norm.weight"])
        _acc(G, p + "attn_layer_norm.weight", dg)
        _acc(G, p + "attn_layer_norm.bias", db)
        dx = d if j == 0 else dx1 + d
    return dx


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
def round_f16(a):
    """Round-to-nearest-even to IEEE half precision (returned in a's dtype)."""
    return np.asarray(a).astype(np.float16).astype(np.asarray(a).dtype)


def forward(P, x, noise, num_iterations, heads, eps=1e-8, folded=True, keep=False, token_dtype=None, weight_dtype=None,
            drop_masks=None):
    """x [B,T,N,D], noise [B,K,Ds] -> slots [B,T,K,Ds], attn [B,T,N,K] (pre-eps softmax).

    dtype follows x (run float64 for the oracle, float32 to measure rounding).
    keep=True also returns everything backward() needs.
    token_dtype="bf16" models the ONE quantisation point of the CUDA bf16 mode: the
    LayerNorm'd token stream xhat is stored in bfloat16 (everything on the slot side
    stays fp32 there, fp64 here); backward() then differentiates that same function
    (rounding treated as identity), which is what the kernels compute.
    weight_dtype="f16" additionally models the tcgen05 clip kernels' weight operands
    (focus_b200/csrc/savi_layout.h: WImg): every matrix a slot-side product multiplies by is
    an fp16 image — the folded products wqk = Ds^-1/2 Wk^T Wq and wg = W_ih W_v (formed in
    full precision, then rounded), W_hh, the MLP and the predictor matrices; biases,
    LayerNorm affines and all activations stay full precision.  (folded form only.)
    drop_masks=(att [Sp,B,H,K,K], out [Sp,B,K,Ds], ffn [Sp,B,K,Ds]) are the predictor's training-mode dropout masks, block
    evaluation f = j*(T-1)+t (the layout of include/focus_savi.h); None = evaluation.
    """
    dt = x.dtype
    P = {k: np.asarray(v, dtype=dt) for k, v in P.items()}
    B, T, N, D = x.shape
    K, Ds = noise.shape[1], noise.shape[2]
    I = num_iterations
    blocks = sum(1 for k in P if k.endswith("attn.proj_q.weight"))
    sc = dt.type(Ds ** -0.5)
    eps = dt.type(eps)

    w16 = weight_dtype == "f16"
    if weight_dtype not in (None, "f16"):
        raise ValueError("weight_dtype must be None or 'f16'")
    if w16:
        assert folded, "weight_dtype models the folded (CUDA) operation order"
        wqk16 = round_f16(sc * (P["project_k.weight"].T @ P["project_q.weight"]))   # [D,Ds]: qk = s~ wqk^T
        wg16 = round_f16(P["gru.weight_ih"] @ P["project_v.weight"])                 # [3Ds,D]: gi = Ux wg^T + b_ih
        whh16 = round_f16(P["gru.weight_hh"])
        w1_16, w2_16 = round_f16(P["mlp.0.weight"]), round_f16(P["mlp.2.weight"])
        Ppred = {k: (round_f16(v) if (k.startswith("predictor.blocks.") and v.ndim == 2) else v) for k, v in P.items()}
    else:
        Ppred = P
    h = P["slot_mu"] + np.exp(P["slot_log_sigma"]) * noise.astype(dt)       # steve.py:56-57
    xhat, zx, rx = _ln(x, P["norm_inputs.weight"], P["norm_inputs.bias"])   # steve.py:60
    if token_dtype == "bf16":
        xhat = round_bf16(xhat)
    if not folded:
        k = (xhat @ P["project_k.weight"].T) * sc                           # steve.py:61,63
        v = xhat @ P["project_v.weight"].T                                  # steve.py:62
    slots_out = np.empty((B, T, K, Ds), dt)
    attn_out = np.empty((B, T, N, K), dt)
    steps, preds = [], []
    for t in range(T):
        for i in range(I):
            hp = h
            st, zs, rs = _ln(hp, P["norm_slots.weight"], P["norm_slots.bias"])      # :72
            q = st @ P["project_q.weight"].T                                         # :75
            if folded:
                qk = st @ wqk16.T if w16 else (q * sc) @ P["project_k.weight"]       # [B,K,D]
                L = xhat[:, t] @ qk.transpose(0, 2, 1)
            else:
                qk = None
                L = k[:, t] @ q.transpose(0, 2, 1)                                   # :76
            Pm = _softmax(L)                                                         # :77
            A = Pm + eps                                                             # :81
            S = A.sum(1)                                                             # [B,K]
            if folded:
                Ux = np.einsum("bnk,bnd->bkd", A, xhat[:, t]) / S[..., None]
                U = Ux @ P["project_v.weight"].T
            else:
                Ux = None
                U = np.einsum("bnk,bnd->bkd", A / S[:, None, :], v[:, t])            # :82-83
            if w16:
                gi = Ux @ wg16.T + P["gru.bias_ih"]
                gh = hp @ whh16.T + P["gru.bias_hh"]
            else:
                gi = U @ P["gru.weight_ih"].T + P["gru.bias_ih"]                     # :87-89
                gh = hp @ P["gru.weight_hh"].T + P["gru.bias_hh"]
            r = _sigmoid(gi[..., :Ds] + gh[..., :Ds])
            z = _sigmoid(gi[..., Ds:2 * Ds] + gh[..., Ds:2 * Ds])
            ghn = gh[..., 2 * Ds:]
            n = np.tanh(gi[..., 2 * Ds:] + r * ghn)
            hg = (1 - z) * n + z * hp
            rec = dict(hp=hp, zs=zs, rs=rs, st=st, q=q, qk=qk, S=S, Ux=Ux, U=U, r=r, z=z, n=n, ghn=ghn, hg=hg, t=t)
            if i < I - 1:                                                            # :92-93
                m, zm, rm = _ln(hg, P["norm_mlp.weight"], P["norm_mlp.bias"])
                a = np.maximum(m @ (w1_16 if w16 else P["mlp.0.weight"]).T + P["mlp.0.bias"], 0)
                h = hg + a @ (w2_16 if w16 else P["mlp.2.weight"]).T + P["mlp.2.bias"]
                rec.update(m=m, zm=zm, rm=rm, a=a)
            else:
                h = hg
            steps.append(rec)
        slots_out[:, t] = h                                                          # :96-97
        attn_out[:, t] = Pm
        if t < T - 1:     # the reference also runs it after the last frame and discards it (:100)
            drop = None
            if drop_masks is not None:
                drop = [tuple(np.asarray(m[j * (T - 1) + t], dt) for m in drop_masks) for j in range(blocks)]
            h, psv = _predictor_fwd(Ppred, h, blocks, heads, drop)
            preds.append(psv)
    if not keep:
        return slots_out, attn_out
    saved = dict(steps=steps, preds=preds, xhat=xhat, zx=zx, rx=rx, noise=noise.astype(dt),
                 I=I, heads=heads, blocks=blocks, eps=eps, shape=(B, T, N, D, K, Ds))
    return slots_out, attn_out, saved


# ----------------------------------------------------------------------------
# backward (folded form; what the CUDA backward kernels implement)
# ----------------------------------------------------------------------------
def backward(P, saved, g_slots, g_attn=None, bwd_weight_dtype=None):
    """Returns (d_inputs [B,T,N,D], {param name: grad}, d_noise [B,K,Ds]).

    bwd_weight_dtype="bf16" models the tcgen05 backward kernel's weight operands (savi_layout.h: WImg, backward
    orientation): every dX = dY . W product multiplies by a bf16 image of W (the folded wg = W_ih W_v and
    wqk = Ds^-1/2 Wk^T Wq are rounded after folding); the weight GRADIENTS dW = dY^T X use no weights and stay exact."""
    B, T, N, D, K, Ds = saved["shape"]
    dt = saved["xhat"].dtype
    P = {k: np.asarray(v, dtype=dt) for k, v in P.items()}
    I, heads, blocks, eps = saved["I"], saved["heads"], saved["blocks"], saved["eps"]
    sc = dt.type(Ds ** -0.5)
    xhat = saved["xhat"]
    G = {}
    dxhat = np.zeros_like(xhat)
    dh = np.zeros((B, K, Ds), dt)
    Wq, Wk, Wv = P["project_q.weight"], P["project_k.weight"], P["project_v.weight"]
    if bwd_weight_dtype not in (None, "bf16"):
        raise ValueError("bwd_weight_dtype must be None or 'bf16'")
    wb = bwd_weight_dtype == "bf16"
    rb = round_bf16 if wb else (lambda a: a)
    W2b, W1b, Whhb = rb(P["mlp.2.weight"]) if "mlp.2.weight" in P else None, rb(P["mlp.0.weight"]), rb(P["gru.weight_hh"])
    wgb = rb(P["gru.weight_ih"] @ Wv)                    # [3Ds, D]
    wqkb = rb(sc * (Wk.T @ Wq))                          # [D, Ds]
    Pp = {k: (rb(v) if (wb and k.startswith("predictor.blocks.") and v.ndim == 2) else v) for k, v in P.items()}
    for t in reversed(range(T)):
        if t < T - 1:
            dh = _predictor_bwd(Pp, saved["preds"][t], dh, blocks, heads, G)
        dh = dh + g_slots[:, t].astype(dt)
        for i in reversed(range(I)):
            s = saved["steps"][t * I + i]
            hp = s["hp"]
            if i < I - 1:
                _acc(G, "mlp.2.bias", dh.sum((0, 1)))
                _acc(G, "mlp.2.weight", np.einsum("bko,bkc->oc", dh, s["a"]))
                da = (dh @ W2b) * (s["a"] > 0)
                _acc(G, "mlp.0.bias", da.sum((0, 1)))
                _acc(G, "mlp.0.weight", np.einsum("bko,bkc->oc", da, s["m"]))
                d, dg, db = _ln_bwd(da @ W1b, s["zm"], s["rm"], P["norm_mlp.weight"])
                _acc(G, "norm_mlp.weight", dg)
                _acc(G, "norm_mlp.bias", db)
                dhg = dh + d
            else:
                dhg = dh
            r, z, n, ghn = s["r"], s["z"], s["n"], s["ghn"]
            dn_pre = dhg * (1 - z) * (1 - n * n)
            dz_pre = dhg * (hp - n) * z * (1 - z)
            dr_pre = dn_pre * ghn * r * (1 - r)
            dgi = np.concatenate([dr_pre, dz_pre, dn_pre], -1)
            dgh = np.concatenate([dr_pre, dz_pre, dn_pre * r], -1)
            dhp = dhg * z + dgh @ Whhb
            _acc(G, "gru.weight_ih", np.einsum("bko,bkc->oc", dgi, s["U"]))
            _acc(G, "gru.weight_hh", np.einsum("bko,bkc->oc", dgh, hp))
            _acc(G, "gru.bias_ih", dgi.sum((0, 1)))
            _acc(G, "gru.bias_hh", dgh.sum((0, 1)))
            dU = dgi @ P["gru.weight_ih"]
            # ---- attention step backward (token pass) ----
            _acc(G, "project_v.weight", np.einsum("bko,bkc->oc", dU, s["Ux"]))
            dUx = dgi @ wgb if wb else dU @ Wv                              # [B,K,D]
            c = (dUx * s["Ux"]).sum(-1)                                     # [B,K]
            xt = xhat[:, t]
            L = xt @ s["qk"].transpose(0, 2, 1)
            Pm = _softmax(L)
            A = Pm + eps
            Sinv = 1.0 / s["S"]
            dWgt = xt @ dUx.transpose(0, 2, 1)                              # [B,N,K]
            dP = (dWgt - c[:, None, :]) * Sinv[:, None, :]
            if g_attn is not None and i == I - 1:
                dP = dP + g_attn[:, t].astype(dt)
            dL = Pm * (dP - (Pm * dP).sum(-1, keepdims=True))
            dqk = np.einsum("bnk,bnd->bkd", dL, xt)
            dxhat[:, t] += dL @ s["qk"] + (A * Sinv[:, None, :]) @ dUx
            # ---- slot-side projections ----
            qs = s["q"] * sc
            _acc(G, "project_k.weight", np.einsum("bko,bkc->oc", qs, dqk))
            dq = (dqk @ Wk.T) * sc
            _acc(G, "project_q.weight", np.einsum("bko,bkc->oc", dq, s["st"]))
            d, dg, db = _ln_bwd(dqk @ wqkb if wb else dq @ Wq, s["zs"], s["rs"], P["norm_slots.weight"])
            _acc(G, "norm_slots.weight", dg)
            _acc(G, "norm_slots.bias", db)
            dh = dhp + d
    # slot initialisation (steve.py:56-57)
    G["slot_mu"] = dh.sum((0, 1)).reshape(1, 1, Ds)
    G["slot_log_sigma"] = (dh * np.exp(P["slot_log_sigma"]) * saved["noise"]).sum((0, 1)).reshape(1, 1, Ds)
    dnoise = dh * np.exp(P["slot_log_sigma"])
    # token LayerNorm (steve.py:60)
    dx, dg, db = _ln_bwd(dxhat, saved["zx"], saved["rx"], P["norm_inputs.weight"])
    G["norm_inputs.weight"] = dg
    G["norm_inputs.bias"] = db
    for name, shp in param_shapes(K, D, Ds, P["mlp.0.weight"].shape[0], blocks).items():
        if name not in G:                      # unused parameters (I == 1, T == 1): zero, never None
            G[name] = np.zeros(shp, dt)
    return dx, G, dnoise


def max_norm_err(a, b):
    """SURVEY.md §8c error metric: max|a-b| / max|b|."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
