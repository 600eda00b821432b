"""torch-CPU restatement of FOCUS's SlotAttentionVideo forward (autograd gives the backward).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this file; nothing under
focus_b200/ does.  It exists because the reference is a Python package that
cannot travel to the GPU box (/root/reference is absent there): this port keeps
the reference's operation order (LayerNorm -> two Linear -> scale -> bmm ->
softmax -> +eps -> token renormalisation -> bmm -> GRUCell -> residual MLP ->
transformer predictor), so timing it on the box's host cores is the reported
"reference CPU" baseline (kind "port"), and autograd through it is a second,
independent gradient oracle next to oracle/savi_numpy.py.

Follows /root/reference/slowfast/models/STEVE/steve.py:52-105 and
transformer.py:22-49, 70-86, 106-114.  Pinned against the reference-generated
fixtures in tests/golden/ by tests/test_oracle_golden.py.
"""
import torch
import torch.nn.functional as F


def _mha(P, pre, y, heads):
    B, K, Ds = y.shape
    dh = Ds // heads
    split = lambda a: a.view(B, K, heads, dh).transpose(1, 2)
    q = split(F.linear(y, P[pre + "proj_q.weight"])) * dh ** -0.5
    k = split(F.linear(y, P[pre + "proj_k.weight"]))
    v = split(F.linear(y, P[pre + "proj_v.weight"]))
    att = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, K, Ds)
    return F.linear(o, P[pre + "proj_o.weight"])


def predictor(P, x, heads):
    Ds = x.shape[-1]
    j = 0
    while ("predictor.blocks.%d.attn.proj_q.weight" % j) in P:
        p = "predictor.blocks.%d." % j
        y = F.layer_norm(x, (Ds,), P[p + "attn_layer_norm.weight"], P[p + "attn_layer_norm.bias"])
        x = (y if j == 0 else x) + _mha(P, p + "attn.", y, heads)
        l2 = F.layer_norm(x, (Ds,), P[p + "ffn_layer_norm.weight"], P[p + "ffn_layer_norm.bias"])
        x = x + F.linear(F.relu(F.linear(l2, P[p + "ffn.0.weight"], P[p + "ffn.0.bias"])),
                         P[p + "ffn.2.weight"], P[p + "ffn.2.bias"])
        j += 1
    return F.layer_norm(x, (Ds,), P["predictor.layer_norm.weight"], P["predictor.layer_norm.bias"])


def forward(P, x, noise, num_iterations, heads, eps=1e-8):
    """P: {reference state_dict name: tensor}; x [B,T,N,D]; noise [B,K,Ds]."""
    B, T, N, D = x.shape
    K, Ds = noise.shape[1], noise.shape[2]
    h = P["slot_mu"] + torch.exp(P["slot_log_sigma"]) * noise
    xh = F.layer_norm(x, (D,), P["norm_inputs.weight"], P["norm_inputs.bias"])
    k = F.linear(xh, P["project_k.weight"]) * Ds ** -0.5
    v = F.linear(xh, P["project_v.weight"])
    slots_out, attn_out = [], []
    for t in range(T):
        kt, vt = k[:, t], v[:, t]
        for i in range(num_iterations):
            hp = h
            q = F.linear(F.layer_norm(h, (Ds,), P["norm_slots.weight"], P["norm_slots.bias"]),
                         P["project_q.weight"])
            pm = torch.softmax(torch.bmm(kt, q.transpose(1, 2)), dim=-1)
            a = pm + eps
            a = a / a.sum(dim=1, keepdim=True)
            u = torch.bmm(a.transpose(1, 2), vt)
            gi = F.linear(u, P["gru.weight_ih"], P["gru.bias_ih"])
            gh = F.linear(hp, P["gru.weight_hh"], P["gru.bias_hh"])
            r = torch.sigmoid(gi[..., :Ds] + gh[..., :Ds])
            z = torch.sigmoid(gi[..., Ds:2 * Ds] + gh[..., Ds:2 * Ds])
            n = torch.tanh(gi[..., 2 * Ds:] + r * gh[..., 2 * Ds:])
            h = (1 - z) * n + z * hp
            if i < num_iterations - 1:
                m = F.layer_norm(h, (Ds,), P["norm_mlp.weight"], P["norm_mlp.bias"])
                h = h + F.linear(F.relu(F.linear(m, P["mlp.0.weight"], P["mlp.0.bias"])),
                                 P["mlp.2.weight"], P["mlp.2.bias"])
        slots_out.append(h)
        attn_out.append(pm)
        # The reference also evaluates the predictor after the last frame and drops
        # the result (steve.py:100); it is kept here so the timed work is the same.
        h = predictor(P, h, heads)
    return torch.stack(slots_out, 1), torch.stack(attn_out, 1)


def forward_backward(P, x, noise, num_iterations, heads, g_slots, g_attn=None, eps=1e-8):
    """Returns slots, attn, d_inputs, {name: grad} via autograd (grads of unused params = 0)."""
    P = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    x = x.detach().clone().requires_grad_(True)
    slots, attn = forward(P, x, noise, num_iterations, heads, eps)
    loss = (slots * g_slots).sum()
    if g_attn is not None:
        loss = loss + (attn * g_attn).sum()
    loss.backward()
    G = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()}
    return slots.detach(), attn.detach(), x.grad, G
