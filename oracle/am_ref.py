"""float64 reference of the acoustic-model arch DSL — TEST INFRASTRUCTURE ONLY (never imported by the product).

Restates, in plain torch float64, the module semantics of the reference's arch files
(recipes/joint_training_vox_populi/cpc/SequentialBuilder.cpp:92-626) for every opcode the four BASELINE archs use
(V RO PD C2 R DO LN TDS L WN C GLU SAUG), so that the CUDA path can be checked end to end — emissions, loss and
every parameter gradient — on the arch files themselves (tests/test_gpu_archs.py).  Parameters are taken from the
trainer's flat arena in module order (layout = wav2letter_b200.trainer.Trainer.layout()):

  C2 / TDS conv  w [kw,1,cin,cout] column-major == memory [cout][cin][kw];  b [cout]
  LN             gain [1], bias [1]                (scalar affine, flashlight LayerNorm with axisSize = -1)
  L / TDS fc     W memory [nOut][nIn] (feature index c*W + w of the internal [B][T][C][W] layout);  b [nOut]
  WN x           v (layout of x's weight), g [nOut], then x's bias
  TDS            conv w,b; LN1 g,b; lin1 W,b; lin2 W,b; LN2 g,b   (tools/StreamingTDSModelConverter.cpp:110-135)

Dropout must be 0 (or the net in eval mode) — masks are not reproducible across implementations; SAUG is the identity
here (tests run it with zero masks or in eval mode).  Activations: TDS family [B,T,C,W]; conv_glu family [B,T,C].
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def same_pad_strided(T: int, k: int, s: int) -> int:
    """flashlight Conv2D PaddingMode::SAME (symmetric), as fl_compat's Conv2D"""
    rem = T % s
    tot = (k - 1) - (s if rem == 0 else rem) + 1
    return max((tot + 1) // 2, 0)


def zero_dropout(arch_text: str) -> str:
    """the arch with every dropout probability set to 0 and SpecAugment's mask counts set to 0 (deterministic forward)"""
    out = []
    for line in arch_text.splitlines():
        p = line.split("#")[0].split()
        if p and p[0] == "DO":
            p[1] = "0.0"
        elif p and p[0] == "TDS" and len(p) > 4:
            p[4] = "0.0"
        elif p and p[0] == "SAUG":
            p[3], p[6] = "0", "0"
        out.append(" ".join(p))
    return "\n".join(out) + "\n"


def parse(arch_text: str, n_feat: int, n_label: int):
    ops = []
    for line in arch_text.splitlines():
        p = line.split("#")[0].replace("NFEAT", str(n_feat)).replace("NLABEL", str(n_label)).split()
        if p:
            ops.append(p)
    return ops


def param_shapes(arch_text: str, n_feat: int, n_label: int):
    """[(elements, fan_in or None)] of every parameter in module order (None: LayerNorm gain / bias, WeightNorm g)"""
    out = []
    for p in parse(arch_text, n_feat, n_label):
        op = p[0]
        if op == "C2":
            cin, cout, k = int(p[1]), int(p[2]), int(p[3])
            out += [(cout * cin * k, cin * k), (cout, cin * k)]
        elif op == "LN":
            out += [(1, "one"), (1, "zero")]
        elif op == "TDS":
            c, k, w = int(p[1]), int(p[2]), int(p[3])
            inner = int(p[5]) if len(p) > 5 and int(p[5]) > 0 else c * w
            out += [(c * c * k, c * k), (c, c * k), (1, "one"), (1, "zero"), (inner * c * w, c * w), (inner, c * w), (c * w * inner, inner),
                    (c * w, inner), (1, "one"), (1, "zero")]
        elif op == "L":
            nin, nout = int(p[1]), int(p[2])
            out += [(nout * nin, nin), (nout, nin)]
        elif op == "WN":
            if p[2] in ("C", "C1"):
                cin, cout, k = int(p[3]), int(p[4]), int(p[5])
                out += [(cout * cin * k, cin * k), (cout, "wn_g"), (cout, cin * k)]
            else:
                nin, nout = int(p[3]), int(p[4])
                out += [(nout * nin, nin), (nout, "wn_g"), (nout, nin)]
    return out


def init_params(arch_text: str, n_feat: int, n_label: int, seed: int = 0, dtype=torch.float32):
    """parameters in module order with flashlight's default init family (uniform, bound sqrt(1 / fan_in)); WeightNorm g =
    the row norms of v so that the wrapped layer is unchanged at initialisation"""
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(arch_text, n_feat, n_label)
    params = []
    for n, fan in shapes:
        if fan == "one":
            params.append(torch.ones(n, dtype=dtype))
        elif fan == "zero":
            params.append(torch.zeros(n, dtype=dtype))
        elif fan == "wn_g":
            v = params[-1].view(n, -1)
            params.append(v.norm(dim=1).clone())
        else:
            b = (1.0 / fan) ** 0.5
            params.append((torch.rand(n, generator=g, dtype=dtype) * 2 - 1) * b)
    return params


class RefNet:
    def __init__(self, arch_text: str, n_feat: int, n_label: int, flat: torch.Tensor = None, layout=None, device=None, params=None,
                 dtype=torch.float64):
        """flat + layout: the trainer's parameter arena (any float dtype) and [(offset, elements, dims)]; or params: a list of
        flat tensors in module order (init_params)"""
        self.ops = parse(arch_text, n_feat, n_label)
        self.dtype = dtype
        if params is not None:
            self.params = [p.detach().to(device or p.device).to(dtype).clone().requires_grad_(True) for p in params]
        else:
            dev = device or flat.device
            self.params = [flat[o:o + n].detach().to(dev).to(dtype).clone().requires_grad_(True) for o, n, _ in layout]

        # conditioning of the scalar LayerNorm parameters, filled by backward(): parameter index -> sum of the ABSOLUTE
        # contributions to its gradient (gain: sum |dy * xhat|, bias: sum |dy|).  These gradients are sums with heavy
        # cancellation; a backward-error bound compares their error with this scale, not with the (tiny) result.
        self.cond = {}
        self._pidx = 0

    # ---- pieces ------------------------------------------------------------------------------------------
    def _ln(self, x, g, b, dims):
        n = F.layer_norm(x, dims, eps=1e-5)
        out = n * g + b
        ig, ib = self._pidx - 2, self._pidx - 1  # gain and bias were the last two parameters drawn
        if out.requires_grad:
            nd = n.detach()

            def note(gr, nd=nd, ig=ig, ib=ib):
                self.cond[ig] = float((gr * nd).abs().sum())
                self.cond[ib] = float(gr.abs().sum())

            out.register_hook(note)
        return out

    def _ln_whole(self, x, g, b):
        return self._ln(x, g, b, x.shape[1:])

    def _ln_frame(self, x, g, b):  # x [B,T,C,W]: normalise every frame over (C, W)
        return self._ln(x, g, b, x.shape[2:])

    @staticmethod
    def _conv_time(x, w, b, stride, pl, pr):
        """x [B,T,C,W]; w [cout,cin,k]"""
        xin = F.pad(x.permute(0, 2, 1, 3), (0, 0, pl, pr))
        return F.conv2d(xin, w.unsqueeze(-1), b, stride=(stride, 1)).permute(0, 2, 1, 3)

    def forward(self, feat: torch.Tensor) -> torch.Tensor:
        """feat [B,1,F,T] (== ArrayFire [T,F,1,B]) -> emissions [B,T',N]"""
        it = iter(self.params)
        self._pidx = 0

        def P():
            self._pidx += 1
            return next(it)

        x = feat.to(self.dtype)
        mode = None  # "tds": [B,T,C,W]; "glu": [B,T,C]; "flat": [B,T,K]
        pend_pad = None
        i = 0
        ops = self.ops
        while i < len(ops):
            p = ops[i]
            op = p[0]
            if op == "V":
                if mode is None:  # head view
                    if p[3] == "1":  # V -1 NFEAT 1 0 -> [T,F,1,B]: W = F, C = 1
                        x = x.permute(0, 3, 1, 2).contiguous()  # [B,T,1,F]
                        mode = "tds"
                    else:  # V -1 1 NFEAT 0 -> [T,1,F,B]: features are channels
                        x = x[:, 0].permute(0, 2, 1).contiguous()  # [B,T,F]
                        mode = "glu"
                elif mode == "tds":  # flatten to the Linear head: feature index c*W + w (fl_compat's convention)
                    B, T, C, W = x.shape
                    x = x.reshape(B, T, C * W)
                    mode = "flat"
                # a trailing `V NLABEL 0 -1 1` is a relabelling
            elif op == "RO":
                pass  # relabelling in the [B,T,...] representation
            elif op == "SAUG":
                pass
            elif op == "PD":
                pend_pad = (int(p[2]), int(p[3]))
            elif op == "C2":
                cin, cout, k, s = int(p[1]), int(p[2]), int(p[3]), int(p[5])
                px = int(p[7]) if len(p) > 7 else 0
                w, b = P().view(cout, cin, k), P().view(cout)
                if pend_pad is not None:
                    pl, pr = pend_pad
                    pend_pad = None
                elif px == -1:
                    pl = pr = same_pad_strided(x.shape[1], k, s)
                else:
                    pl = pr = px
                x = self._conv_time(x, w, b, s, pl, pr)
            elif op == "R":
                x = x.clamp_min(0)
            elif op == "DO":
                if float(p[1]) != 0.0:
                    raise ValueError("reference: dropout must be 0 (use zero_dropout())")
            elif op == "LN":
                axes = sorted(int(a) for a in p[1:])
                g, b = P(), P()
                x = self._ln_frame(x, g, b) if axes == [1, 2] else self._ln_whole(x, g, b)
            elif op == "TDS":
                c, k, w_ = int(p[1]), int(p[2]), int(p[3])
                if len(p) > 4 and float(p[4]) != 0.0:
                    raise ValueError("reference: dropout must be 0 (use zero_dropout())")
                inner = int(p[5]) if len(p) > 5 and int(p[5]) > 0 else c * w_
                rpad = int(p[6]) if len(p) > 6 else -1
                ln_time = (int(p[7]) != 0) if len(p) > 7 else True
                ln = self._ln_whole if ln_time else self._ln_frame
                cw, cb = P().view(c, c, k), P().view(c)
                if rpad < 0:
                    pl = pr = same_pad_strided(x.shape[1], k, 1)
                else:
                    pl, pr = k - 1 - rpad, rpad
                y1 = self._conv_time(x, cw, cb, 1, pl, pr).clamp_min(0)
                g1, b1 = P(), P()
                z = ln(x + y1, g1, b1)  # (called right after its two parameters are drawn: _ln notes their indices)
                W1, bb1, W2, bb2 = P().view(inner, c * w_), P(), P().view(c * w_, inner), P()
                B, T = z.shape[:2]
                f = z.reshape(B, T, c * w_)
                u = F.linear(F.linear(f, W1, bb1).clamp_min(0), W2, bb2)
                g2, b2 = P(), P()
                x = ln(z + u.view(B, T, c, w_), g2, b2)
            elif op == "L":
                nin, nout = int(p[1]), int(p[2])
                W, b = P().view(nout, nin), (P() if (len(p) < 4 or int(p[3]) != 0) else None)
                if mode == "tds":
                    B, T, C, Wd = x.shape
                    x = x.reshape(B, T, C * Wd)
                    mode = "flat"
                x = F.linear(x, W, b)
            elif op == "WN":
                dim, kind = int(p[1]), p[2]
                if kind in ("C", "C1"):
                    cin, cout, k, s = int(p[3]), int(p[4]), int(p[5]), int(p[6])
                    pad = int(p[7]) if len(p) > 7 else 0
                    assert dim == 3 and s == 1
                    v, g, b = P().view(cout, cin * k), P().view(cout, 1), P().view(cout)
                    w = (g * v / v.norm(dim=1, keepdim=True)).view(cout, cin, k)
                    pl = pr = (k // 2) if pad == -1 else pad
                    x = F.conv1d(F.pad(x.permute(0, 2, 1), (pl, pr)), w, b).permute(0, 2, 1)
                else:
                    nin, nout = int(p[3]), int(p[4])
                    assert dim == 0 and kind == "L"
                    v, g, b = P().view(nout, nin), P().view(nout, 1), P().view(nout)
                    x = F.linear(x, g * v / v.norm(dim=1, keepdim=True), b)
            elif op == "GLU":
                h = x.shape[-1] // 2
                x = x[..., :h] * torch.sigmoid(x[..., h:])
            else:
                raise ValueError(f"opcode {op} not covered by the float64 reference")
            i += 1
        rest = list(it)
        if rest:
            raise ValueError(f"{len(rest)} parameters left unconsumed: arch / layout mismatch")
        return x

    def grads_flat(self, layout, total: int) -> torch.Tensor:
        out = torch.zeros(total, dtype=self.dtype, device=self.params[0].device)
        for (o, n, _), p in zip(layout, self.params):
            if p.grad is not None:
                out[o:o + n] = p.grad.reshape(-1)
        return out
