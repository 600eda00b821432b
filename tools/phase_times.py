"""Per-phase cycle breakdown of the clip kernels (CTA 0), via savi_debug_set_phase_buffer, on a -DSAVI_PHASE_PROFILE build of the
library (tools/build/libfocus_savi_phases.so, built on first use: the production library has no probes in its tcgen05 kernels)."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the production library compiles the phase probes of the tcgen05 clip kernels out: build (once) a profiling variant and load THAT
_spec = importlib.util.spec_from_file_location("_focus_b200_build", os.path.join(ROOT, "focus_b200", "build.py"))
_b = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(_b)
_lib_path = os.path.join(ROOT, "tools", "build", "libfocus_savi_phases.so")
if not os.path.exists(_lib_path) or os.path.getmtime(_lib_path) < _b._newest_source_mtime():
    os.makedirs(os.path.dirname(_lib_path), exist_ok=True)
    _b.build(force=True, verbose=False, out=_lib_path, extra_flags=["-DSAVI_PHASE_PROFILE"])
os.environ["FOCUS_SAVI_LIB"] = _lib_path
import torch, bench
from focus_b200 import _lib
NAMES = {0: "top/pred-tail", 1: "copy+LN", 2: "lin q", 3: "lin qk", 4: "token pass", 5: "cluster sync", 6: "combine", 7: "lin U",
         8: "lin gi", 9: "lin gh", 10: "gru pointwise", 11: "mlp LN", 12: "lin a", 13: "lin h(mlp2)", 14: "slots copy", 15: "pred LN1",
         16: "pred QKV", 17: "pred mha core", 18: "pred Wo", 19: "pred LN2", 20: "pred F1", 21: "pred F2",
         30: "B: top", 31: "B pred: LNf..dO", 32: "B pred mha core", 33: "B pred: dy..LN1", 34: "B step setup", 35: "B mlp bwd",
         36: "B gru pointwise+colsum", 37: "B lin dh(whh)", 38: "B lin dU(wih)", 39: "B lin dUx + cvec", 40: "B token pass",
         41: "B cluster sync", 42: "B combine", 43: "B lin dq", 44: "B LN st", 45: "B lin dst", 46: "B LN bwd", 50: "tp: stage qk + issue", 51: "tp: wait tile", 52: "tp: phase1 mma", 53: "tp: softmax", 54: "tp: stores+sync", 55: "tp: attn copy", 56: "tp: phase2", 57: "tp: epilogue", 58: "lin: entry+stage X", 59: "lin: sync", 60: "lin: issue B loads", 61: "lin: wait B + MMAs", 62: "lin: epilogue", 63: "lin: final sync"}
UN = {1: "u: save hp, LN_s, s~ operand", 2: "u: WAIT q", 3: "u: q epilogue", 4: "u: WAIT qk", 5: "u: qk epilogue", 6: "u: softmax tiles (total)",
      7: "u: WAIT token pass", 8: "u: ld numx + send", 9: "u: WAIT peer", 10: "u: Ux epilogue", 11: "u: WAIT U", 12: "u: U epilogue", 13: "u: WAIT gru",
      14: "u: gru math + saves", 15: "u: mlp LN + operand", 16: "u: WAIT a", 17: "u: a epilogue", 18: "u: WAIT h2", 19: "u: h2", 20: "u: predictor (total)",
      50: "  sm: WAIT logits", 51: "  sm: ld + softmax", 52: "  sm: WAIT aw free", 53: "  sm: write A + signal", 54: "  sm: attn out",
      21: "  p: LN1, WAIT q,k,v", 22: "  p: epilogue + attention core", 23: "  p: WAIT proj_o, residual, LN2", 24: "  p: ffn tiles, WAIT ffn.2, final LN"}
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
c = dict(bench.CONFIGS[cfg])
dt = torch.float32 if c["dtype"] == "fp32" else torch.bfloat16
m = bench.make_params_like(c).cuda()
g = torch.Generator().manual_seed(1)
x = torch.randn(c["B"], c["T"], c["N"], c["D"], generator=g).to(dt).cuda().requires_grad_(True)
noise = torch.randn(c["B"], c["K"], c["Ds"], generator=g).cuda()
gs = torch.randn(c["B"], c["T"], c["K"], c["Ds"], generator=g).cuda()
ga = torch.randn(c["B"], c["T"], c["N"], c["K"], generator=g).to(dt).cuda()
def step():
    s, at = m(x, noise=noise)
    torch.autograd.backward([s, at], [gs.to(s.dtype), ga]); x.grad = None
step(); torch.cuda.synchronize()
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
_lib.lib.savi_debug_set_phase_buffer(buf.data_ptr())
step(); torch.cuda.synchronize()
_lib.lib.savi_debug_set_phase_buffer(None)
v = buf.cpu().tolist()
if os.environ.get("SAVI_DISABLE_UMMA") is None and c["D"] == 128:
    tot = sum(v[1:21])
    print("UMMA forward, compute thread 0 of CTA 0: total %.1f us" % (tot / 1965.0))
    for i in sorted(UN):
        if v[i]: print("  %-34s %8.1f us  %5.1f%%" % (UN[i], v[i] / 1965.0, 100.0 * v[i] / tot))
    print("  issuer thread: total %.1f us | waiting for operands %.1f | waiting for ring blocks %.1f | token pass (incl. its waits) %.1f" % tuple(v[i] / 1965.0 for i in (63, 60, 61, 62)))
    for i in range(25): v[i] = 0
    for i in range(50, 55): v[i] = 0
    for i in range(60, 64): v[i] = 0
    BN = {25: "b: predictor bwd + grad_slots (per frame)", 26: "b: mlp bwd", 27: "b: gru bwd + operands", 28: "b: loads, 1/S", 29: "b: WAIT dUx",
          30: "b: c vector + operands", 31: "b: softmax-bwd tiles (total)", 32: "b: WAIT token pass", 33: "b: dqk + exchange", 34: "b: dqk operand",
          35: "b: WAIT ds~", 36: "b: LN_s bwd", 37: "  pb: LN_f / LN1 bwd (+ frame top)", 38: "  pb: dx2 operand, 4 ffn.2^T tiles, df chunks", 39: "  pb: WAIT d l2", 40: "  pb: LN2 bwd + dx1 operand", 41: "  pb: loads + WAIT dO", 42: "  pb: mha core bwd", 43: "  pb: saves + 3 operands", 44: "  pb: WAIT dy", 45: "    core: stage tiles + barrier", 46: "    core: d attention", 47: "    core: softmax bwd", 48: "    core: dQ dK dV",
          55: "  sb: WAIT logits (incl. grad_attn load)", 56: "  sb: ld + softmax + dP + dot", 57: "  sb: WAIT dl free",
          58: "  sb: write dL + signal", 59: "  sb: coef stores"}
    tot = sum(v[25:37])
    if tot:
        print("UMMA backward, compute thread 0 of CTA 0: total %.1f us" % (tot / 1965.0))
        for i in sorted(BN):
            if v[i]: print("  %-42s %8.1f us  %5.1f%%" % (BN[i], v[i] / 1965.0, 100.0 * v[i] / tot))
        for i in list(range(25, 50)) + list(range(55, 60)): v[i] = 0
tot_f = sum(v[:30]); tot_b = sum(v[30:50])
print("forward  total %.1f us (cycles @1.965GHz)" % (tot_f / 1965.0))
for i in range(30):
    if v[i]: print("  %-24s %8.1f us  %5.1f%%" % (NAMES.get(i, i), v[i] / 1965.0, 100.0 * v[i] / tot_f))
print("backward total %.1f us" % (tot_b / 1965.0))
for i in range(30, 50):
    if v[i]: print("  %-24s %8.1f us  %5.1f%%" % (NAMES.get(i, i), v[i] / 1965.0, 100.0 * v[i] / tot_b))
print('token pass fwd sub-phases (included in "token pass")')
for i in range(50, 64):
    if v[i]: print("  %-24s %8.1f us" % (NAMES.get(i, i), v[i] / 1965.0))
