"""Per-point check of the gathering colour head (k3_color_ws.cu) against round 1's record-fed one on one frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import gpnerf_b200  # noqa
from gpnerf_b200 import synth
from gpnerf_b200._lib import PREC_BF16
import stages
H = int(os.environ.get("H", "72")); V = int(os.environ.get("V", "3")); S = int(os.environ.get("S", "40"))
scene = synth.make_scene("zju", H=H, W=H, V=V, seed=81)
w = synth.make_head_weights(V=V, seed=141, random_bias=True)
os.environ["GPNERF_COLOR_IMPL"] = "records"
e0, _ = stages.run_engine_progressive(scene, w, S, precision=PREC_BF16)
os.environ["GPNERF_COLOR_IMPL"] = ""
e1, _ = stages.run_engine_progressive(scene, w, S, precision=PREC_BF16)
c0, c1 = e0.read_counters(), e1.read_counters()
print(c0, c1)
p1, p2 = c1["P1"], c1["P2"]
print("sigma equal", torch.equal(e0.sigma[:p1], e1.sigma[:p1]), "valid1 equal", torch.equal(e0.valid1[:p2], e1.valid1[:p2]))
v1 = e1.valid1[:p2].long()
r0 = e0.rgb[: p1 * 3].view(p1, 3)[v1]
r1 = e1.rgb[: p1 * 3].view(p1, 3)[v1]
bad = ~torch.isfinite(r1).all(1)
print("non-finite rows", int(bad.sum()), "of", p2, "first bad positions", torch.nonzero(bad)[:10].flatten().tolist())
d = (r1 - r0).abs()
d[bad] = 0
print("max |d rgb| over finite rows", float(d.max()), "mean", float(d.mean()))
rowmax = d.max(1)[0]
worst = torch.argsort(rowmax, descending=True)[:10]
print("worst rows (position in valid1, tile row):", [(int(i), int(i) % 128, round(float(rowmax[i]), 4)) for i in worst])
print("image diff", float((e1.pred_img - e0.pred_img).abs().nan_to_num(9.0).max()))
import ctypes as C
lib = e1.lib
if hasattr(lib, "gpnerf_debug_color"):
    import struct
    buf = (C.c_uint * 304)()
    lib.gpnerf_debug_color(buf)
    b = list(buf)
    print("non-finite counts per stage (acc1..acc6):", b[:6])
    print("smem image vs global image at kernel end: mismatching words", b[6], "first byte", b[7], "last byte", b[8], "of", b[9])
    for code in range(6):
        print("stage", code, "row5 values", [round(struct.unpack("f", struct.pack("I", v))[0], 3) for v in b[16 + code * 16: 32 + code * 16]])
        print("        non-finite per column", b[112 + code * 16: 128 + code * 16])
img = e1._weight_tensors[-1]
bf = img.view(torch.bfloat16).float()
bad = torch.nonzero(~torch.isfinite(bf)).flatten()
print("image bytes", img.numel(), "non-finite bf16 at byte offsets", (bad * 2).tolist()[:40], "count", bad.numel())
img0 = e0._weight_tensors[-1]
print("images equal (records engine vs gathering engine):", torch.equal(img, img0))
# ---- reference pre-activations of survivor t = 5 (row 5 of tile 0) from the exact fp32 engine's gathered rows
if hasattr(lib, "gpnerf_debug_color"):
    import math
    from gpnerf_b200._lib import PREC_FP32
    ef, _ = stages.run_engine_progressive(scene, w, S, precision=PREC_FP32)
    t = 5
    i = int(e1.valid1[t])
    rf = ef.rgb_feat[: p1 * V * 35].view(p1, V, 35)[i].cpu()
    mv = ef.meanvar[: p1 * 70].view(p1, 70)[i].cpu()
    W = {k: v.float() for k, v in w.items()}
    elu = torch.nn.functional.elu
    c = 1.0 / math.log(2.0)
    xs, es, pres = [], [], {}
    for v in range(V):
        inp = torch.cat([mv, rf[v]])
        pre1 = W["rgbhead.base_fc.0.weight"] @ inp + W["rgbhead.base_fc.0.bias"]
        x = elu(W["rgbhead.base_fc.2.weight"] @ elu(pre1) + W["rgbhead.base_fc.2.bias"])
        pre2 = W["rgbhead.base_fc.2.weight"] @ elu(pre1) + W["rgbhead.base_fc.2.bias"]
        pre3 = W["rgbhead.vis_fc.0.weight"] @ (x / V) + W["rgbhead.vis_fc.0.bias"]
        pre4 = W["rgbhead.vis_fc.2.weight"] @ elu(pre3) + W["rgbhead.vis_fc.2.bias"]
        xs.append(x); es.append(elu(pre4))
        pres[v] = (pre1, pre2, pre3, pre4)
    flat = torch.cat([xs[v] + es[v] for v in range(V)])
    pre5 = W["rgbhead.rgb_fc.0.weight"] @ flat + W["rgbhead.rgb_fc.0.bias"]
    pre6 = W["rgbhead.rgb_fc.2.weight"] @ elu(pre5) + W["rgbhead.rgb_fc.2.bias"]
    v = V - 1
    print("expected (last view) acc1[48:64]", [round(float(a) * c, 3) for a in pres[v][0][48:64]])
    print("expected acc2[16:32]", [round(float(a) * c, 3) for a in pres[v][1][16:32]])
    print("expected acc3[16:32]", [round(float(a) * c, 3) for a in pres[v][2][16:32]])
    print("expected acc4[16:32]", [round(float(a) * c, 3) for a in pres[v][3][16:32]])
    print("expected acc5[16:32]", [round(float(a) * c, 3) for a in pre5[16:32]])
    pre5e = W["rgbhead.rgb_fc.0.weight"] @ torch.cat(es) + W["rgbhead.rgb_fc.0.bias"]
    pre5x = W["rgbhead.rgb_fc.0.weight"] @ torch.cat(xs)
    print("expected acc5[16:32] E part + bias only", [round(float(a) * c, 3) for a in pre5e[16:32]])
    print("expected acc5[16:32] x part only", [round(float(a) * c, 3) for a in pre5x[16:32]])
    print("expected acc6[0:16]", [round(float(a) * c, 3) for a in pre6[0:16]])

    def bf16_pairs(words):
        out = []
        for wd in words:
            for half in (wd & 0xffff, wd >> 16):
                out.append(struct.unpack("f", struct.pack("I", half << 16))[0])
        return torch.tensor(out)
    XSd = bf16_pairs(b[208 + 16: 208 + 32])      # code 1: x~/V (32 values)
    Ed = bf16_pairs(b[208 + 48: 208 + 64])       # code 3: e~
    Zd = bf16_pairs(b[208 + 64: 208 + 80])       # code 4: z~
    print("XS dumped", [round(float(a), 3) for a in XSd[:8]], "expected x~/V", [round(float(a) * c / V, 3) for a in xs[V - 1][:8]])
    print("E dumped", [round(float(a), 3) for a in Ed[:8]], "expected e~", [round(float(a) * c, 3) for a in es[V - 1][:8]])
    W0r = W["rgbhead.rgb_fc.0.weight"].bfloat16().float(); b0r = W["rgbhead.rgb_fc.0.bias"]
    if V == 1:
        acc5_from_dump = W0r @ Ed + (W0r * V).bfloat16().float() @ XSd + b0r * c
        print("acc5[16:32] recomputed from the dumped operands", [round(float(a), 3) for a in acc5_from_dump[16:32]])
        W1r = W["rgbhead.rgb_fc.2.weight"].bfloat16().float()
        print("acc6[0:16] recomputed from dumped Z", [round(float(a), 3) for a in (W1r @ Zd + W["rgbhead.rgb_fc.2.bias"] * c)[:16]])
    if V == 1:
        kern = torch.tensor([struct.unpack("f", struct.pack("I", v))[0] for v in b[16 + 4 * 16: 16 + 5 * 16]])
        print("kernel - recomputed", [round(float(a), 3) for a in (kern - acc5_from_dump[16:32])])
        print("c*bias[16:32]", [round(float(a) * c, 3) for a in b0r[16:32]])
        xe = Ed + XSd
        for name, sel in (("k 0..15 only", slice(0, 16)), ("k 16..31 only", slice(16, 32))):
            part = W0r[:, sel] @ xe[sel]
            print("contribution of", name, [round(float(a), 3) for a in part[16:32]])
        Wv1 = W["rgbhead.vis_fc.2.weight"].bfloat16().float()
        print("acc6[0:16] if B were vis_fc.2 rows 0..15", [round(float(a), 3) for a in (Wv1 @ Zd + W["rgbhead.vis_fc.2.bias"] * c)[:16]])
