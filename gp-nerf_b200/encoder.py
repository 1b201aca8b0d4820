"""The image encoder: mirror of libs/encoders/UNet.py:133-234 (`ResUNet`, resnet34 trunk:
7×7/2 stem, 3+4+6 BasicBlocks at strides 2/2/2, two up-convolutions with skip connections, 1×1 head;
reflect padding everywhere, InstanceNorm2d(affine, no running statistics) after every convolution).

Same attribute tree, hence the same ``state_dict`` keys and shapes as the reference (``conv1.weight``,
``layer2.0.downsample.1.bias``, ``upconv3.conv.bn.weight``, ``out_conv.bias`` …): ``encoder.*`` of a reference
checkpoint loads with ``strict=True``.  The nn.Conv2d / nn.InstanceNorm2d children are parameter holders; the
forward pass is its own launch sequence:

* convolutions: cuDNN (library call – SURVEY §8f: "cuDNN first, custom only if it dominates") on channels-last
  activations, fp16 by default (instance-normalised activations stay O(1): fp16's 11-bit mantissa keeps the
  36-layer chain within 1 % of the fp32 result where bf16 drifts to 6 %; bf16 and fp32-without-TF32 selectable);
* every InstanceNorm + ReLU / + residual + ReLU / + ELU: the two K9 kernels (csrc/k9_instnorm.cu) – one
  statistics pass, one fused apply pass that also writes the reflected one-pixel border the next 3×3
  convolution needs (no F.pad pass, no NCHW round trip) and can target a channel slice of a concatenation
  buffer; bilinear ×2 upsampling and the skip copies by `resample_pad` in the same layout;
* the whole sequence is captured into one CUDA graph per input shape.

Inference form (no autograd through the kernels).  Pinned against the reference module's own output on seeded
parameters and inputs: tests/golden/encoder.npz (oracle/gen_golden_encoder.py).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr

ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
_DTYPES = {"fp32": (torch.float32, 0), "bf16": (torch.bfloat16, 1), "fp16": (torch.float16, 2)}


def _conv(c_in, c_out, k, stride=1, bias=False):
    return nn.Conv2d(c_in, c_out, kernel_size=k, stride=stride, padding=(k - 1) // 2, bias=bias, padding_mode="reflect")


def _norm(c):
    return nn.InstanceNorm2d(c, track_running_stats=False, affine=True)


class BasicBlock(nn.Module):
    """UNet.py:16-53 (parameter holder)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = _conv(inplanes, planes, 3, stride), _norm(planes)
        self.conv2, self.bn2 = _conv(planes, planes, 3), _norm(planes)
        self.downsample = downsample
        self.stride = stride


class conv(nn.Module):      # noqa: N801 – the reference's class name (UNet.py:106)
    def __init__(self, num_in_layers, num_out_layers, kernel_size, stride):
        super().__init__()
        self.conv = _conv(num_in_layers, num_out_layers, kernel_size, stride, bias=True)
        self.bn = _norm(num_out_layers)


class upconv(nn.Module):    # noqa: N801 – UNet.py:122
    def __init__(self, num_in_layers, num_out_layers, kernel_size, scale):
        super().__init__()
        self.scale = scale
        self.conv = conv(num_in_layers, num_out_layers, kernel_size, 1)


class ResUNet(nn.Module):
    def __init__(self, encoder="resnet34", out_ch=32, norm_layer=None, precision="fp16", use_cuda_graph=True):
        super().__init__()
        if encoder not in ("resnet18", "resnet34"):
            raise _lib.GpnerfError("ResUNet: the reference builds BasicBlock trunks only (UNet.py:155); "
                                   "resnet18/resnet34 filter widths")
        if precision not in _DTYPES:
            raise _lib.GpnerfError("precision must be 'fp16', 'bf16' or 'fp32'")
        self.precision, self.use_cuda_graph = precision, use_cuda_graph
        self.conv1, self.bn1 = _conv(3, 64, 7, 2), _norm(64)
        self._inplanes = 64
        self.layer1 = self._make_layer(64, 3, 2)          # UNet.py:166-170: layers [3, 4, 6], all stride 2
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.upconv3 = upconv(256, 128, 3, 2)
        self.iconv3 = conv(128 + 128, 128, 3, 1)
        self.upconv2 = upconv(128, 64, 3, 2)
        self.iconv2 = conv(64 + 64, out_ch, 3, 1)
        self.out_conv = nn.Conv2d(out_ch, out_ch, 1, 1)
        self._packed = None
        self._plist = None
        self._scratch = {}           # per device: zeroed once, self-cleaning (csrc/k9_instnorm.cu)
        self._graphs = {}

    def _make_layer(self, planes, blocks, stride):
        down = None
        if stride != 1 or self._inplanes != planes:
            down = nn.Sequential(_conv(self._inplanes, planes, 1, stride), _norm(planes))
        layers = [BasicBlock(self._inplanes, planes, stride, down)]
        self._inplanes = planes
        layers += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ parameters in kernel form
    def _params(self, device):
        """Convolution weights in the compute dtype, channels-last; norm affine parameters fp32.  Refreshed in
        place when a parameter changes (the CUDA graphs keep reading the same memory)."""
        if self._plist is None:                  # walking the module tree costs 0.5 ms per call
            self._plist = list(self.parameters())
        ver = tuple((p.data_ptr(), p._version) for p in self._plist)
        if self._packed is not None and self._packed[0] == ver and self._packed[1] == device:
            return self._packed[2]
        dt = _DTYPES[self.precision][0]
        fresh = {}
        for name, p in self.named_parameters():
            v = p.detach().to(device)
            if v.dim() == 4:
                fresh[name] = v.to(dt).contiguous(memory_format=torch.channels_last)
            elif ".bn" in name or name.startswith("bn") or ".downsample.1." in name:
                fresh[name] = v.float().contiguous()
            else:                                             # convolution biases
                fresh[name] = v.to(dt).contiguous()
        if self._packed is not None and self._packed[1] == device:
            for k, v in fresh.items():
                self._packed[2][k].copy_(v)
            fresh = self._packed[2]
        else:
            self._graphs = {}
        self._packed = (ver, device, fresh)
        return fresh

    # ------------------------------------------------------------------ launch sequence
    # Activations live in channels-last buffers [N, H+2p, W+2p, C]; p = 1 buffers carry their reflected border
    # (written by the producing K9 kernel), so every 3×3 convolution is an unpadded cuDNN call on the buffer.
    def _buf(self, n, h, w, c, pad, device):
        return torch.empty(n, h + 2 * pad, w + 2 * pad, c, dtype=_DTYPES[self.precision][0], device=device)

    @staticmethod
    def _st(device):
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    def _norm_act(self, x, w, prefix, act, residual=None, res_pad=0, out=None, out_pad=0, coff=0):
        """x: convolution output (NCHW view of dense channels-last memory) → `out` buffer (allocated if None)."""
        lib = _lib.load()
        x = x.contiguous(memory_format=torch.channels_last)
        n, c, h, wd = x.shape
        if out is None:
            out = self._buf(n, h, wd, c, out_pad, x.device)
        nc_cap = n * 256                                   # widest layer of the trunk
        sc = self._scratch.get(x.device)
        if sc is None or sc[1] < max(nc_cap, n * c):
            nc_cap = max(nc_cap, n * c)
            sc = self._scratch[x.device] = (torch.zeros(32 + nc_cap * 3, dtype=torch.float64, device=x.device), nc_cap)
        dp = lambda t: None if t is None else C.c_void_p(t.data_ptr())       # noqa: E731
        check(lib.gpnerf_k9_instance_norm_act(dp(x), dp(residual), res_pad, _DTYPES[self.precision][1], n, h, wd, c,
                                              ptr(w[prefix + ".weight"]), ptr(w[prefix + ".bias"]), 1e-5, act,
                                              ptr(sc[0]), sc[1], dp(out), out_pad, out.shape[3], coff, self._st(x.device)),
              "instance_norm_act")
        return out

    def _resample(self, src, src_pad, mode, out=None, out_pad=1, coff=0):
        lib = _lib.load()
        n, hs, ws, c = src.shape[0], src.shape[1] - 2 * src_pad, src.shape[2] - 2 * src_pad, src.shape[3]
        h, wd = {0: (hs, ws), 1: (2 * hs, 2 * ws), 2: ((hs + 1) // 2, (ws + 1) // 2)}[mode]
        if out is None:
            out = self._buf(n, h, wd, c, out_pad, src.device)
        check(lib.gpnerf_k9_resample_pad(C.c_void_p(src.data_ptr()), _DTYPES[self.precision][1], n, hs, ws, src_pad, c, mode,
                                         C.c_void_p(out.data_ptr()), h, wd, out_pad, out.shape[3], coff,
                                         self._st(src.device)), "resample_pad")
        return out

    @staticmethod
    def _conv_p(buf, weight, bias, stride):
        """cuDNN convolution on a bordered buffer (its border is the padding)."""
        return F.conv2d(buf.permute(0, 3, 1, 2), weight, bias, stride)

    def _block(self, a, w, prefix, blk):
        """a: pad-1 buffer → pad-1 buffer (UNet.py:38-53)."""
        y = self._norm_act(self._conv_p(a, w[prefix + ".conv1.weight"], None, blk.stride), w, prefix + ".bn1", ACT_RELU,
                           out_pad=1)
        y = self._conv_p(y, w[prefix + ".conv2.weight"], None, 1)
        if blk.downsample is not None:
            # 1×1 stride-2 convolution = a GEMM over every second pixel
            sub = self._resample(a, 1, 2, out_pad=0) if blk.stride == 2 else a[:, 1:-1, 1:-1, :].contiguous()
            idt = F.conv2d(sub.permute(0, 3, 1, 2), w[prefix + ".downsample.0.weight"])
            idt = self._norm_act(idt, w, prefix + ".downsample.1", ACT_NONE)
            return self._norm_act(y, w, prefix + ".bn2", ACT_RELU, residual=idt, res_pad=0, out_pad=1)
        return self._norm_act(y, w, prefix + ".bn2", ACT_RELU, residual=a, res_pad=1, out_pad=1)

    def _run(self, x, w):
        dt = _DTYPES[self.precision][0]
        n, _, hi, wi = x.shape
        x = F.pad(x.to(dt), (3, 3, 3, 3), mode="reflect").contiguous(memory_format=torch.channels_last)
        a = self._norm_act(F.conv2d(x, w["conv1.weight"], None, 2), w, "bn1", ACT_RELU, out_pad=1)
        feats = []
        for name in ("layer1", "layer2", "layer3"):
            for i, blk in enumerate(getattr(self, name)):
                a = self._block(a, w, f"{name}.{i}", blk)
            feats.append(a)
        x1, x2, x3 = feats

        def up_cat(src, src_pad, skip, prefix):
            """upconv (×2 bilinear → conv → IN → ELU) into the first channels of the concatenation buffer, the
            skip tensor into the rest (UNet.py:122-131, 204-216, 223-229)."""
            u = self._resample(src, src_pad, 1)
            y = self._conv_p(u, w[prefix + ".conv.conv.weight"], None, 1)      # bias: cancelled by the norm
            c_up, c_skip = y.shape[1], skip.shape[3]
            hs, ws = skip.shape[1] - 2, skip.shape[2] - 2
            if (hs, ws) == tuple(y.shape[2:]):
                cat = self._buf(n, hs, ws, c_up + c_skip, 1, y.device)
                self._norm_act(y, w, prefix + ".conv.bn", ACT_ELU, out=cat, out_pad=1, coff=0)
                self._resample(skip, 1, 0, out=cat, out_pad=1, coff=c_up)
                return cat
            # sizes that do not halve evenly: the reference zero-pads the skip tensor (UNet.py:205-209)
            up = self._norm_act(y, w, prefix + ".conv.bn", ACT_ELU).permute(0, 3, 1, 2)
            sk = skip[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
            dy, dx = up.shape[2] - sk.shape[2], up.shape[3] - sk.shape[3]
            sk = F.pad(sk, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
            cat = torch.cat([up, sk], dim=1)
            return F.pad(cat, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1).contiguous()

        cat3 = up_cat(x3, 1, x2, "upconv3")
        # a per-channel bias in front of an InstanceNorm drops out of (x − mean): the four conv biases of the
        # decoder are loaded (state_dict) but never added
        i3 = self._norm_act(self._conv_p(cat3, w["iconv3.conv.weight"], None, 1), w, "iconv3.bn", ACT_ELU)
        cat2 = up_cat(i3, 0, x1, "upconv2")
        i2 = self._norm_act(self._conv_p(cat2, w["iconv2.conv.weight"], None, 1), w, "iconv2.bn", ACT_ELU)
        y = F.conv2d(i2.permute(0, 3, 1, 2), w["out_conv.weight"], w["out_conv.bias"])
        return y.float().contiguous()                           # [V, out_ch, H/4, W/4] fp32 NCHW, as the reference

    @torch.no_grad()
    def forward(self, x):
        if x.device.type != "cuda":
            raise _lib.GpnerfError("gpnerf_b200 runs on CUDA devices only (no CPU fallback)")
        if self._plist is not None and self._plist[0].device != x.device:
            self._plist = None                   # .to(device) replaced the parameters
        w = self._params(x.device)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
                return self._run(x, w)
            key = (tuple(x.shape), x.dtype)
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = {"uses": 0, "graph": None, "x": torch.empty_like(x)}
            g["uses"] += 1
            if g["graph"] is None and g["uses"] < 3:            # cuDNN picks its algorithms on the eager runs
                return self._run(x, w)
            g["x"].copy_(x, non_blocking=True)
            if g["graph"] is None:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    g["y"] = self._run(g["x"], w)
                g["graph"] = graph
            g["graph"].replay()
            return g["y"]


def build_encoder(cfg):
    """UNet.py:237-243"""
    return ResUNet(encoder=cfg.encoder.name, out_ch=cfg.encoder.out_ch)
