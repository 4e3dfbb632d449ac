"""Test infrastructure: a CPU emulation of the CONTRACT of wm_pconv_fwd / wm_pconv_in_fwd / wm_pconv_to_planar /
wm_pconv_from_planar (include/wmb200.h) on fp32 tensors, used to check pconv.py's translation of the reference's layers
into GEMM descriptions without a GPU.  Planar tensors are [phase][C][plane_rows] fp32, initialised to NaN so that any read
of a row nobody wrote (or of memory in front of / behind a plane) poisons the result the test compares."""
import torch
import torch.nn.functional as F

from wmb200 import pconv as PC

GAP = PC.GAP
FRONT = 8


class EmuPlanar:
    def __init__(self, Cn, B, T, split=1):
        self.C, self.B, self.T, self.split = Cn, B, T, split
        self.RP = PC.plane_rows(B, T)
        self.phase_rows = 2 * (Cn // 8) * self.RP
        self.store = torch.full((split, Cn, self.RP), float("nan"), dtype=torch.float64)


class EmuBackend:
    def __init__(self):
        self.calls = []

    def planar(self, Cn, B, T, split, device):
        return EmuPlanar(Cn, B, T, split)

    def fp32(self, shape, device):
        return torch.full(shape, float("nan"), dtype=torch.float64)

    def conv_in(self, s, conv, out):
        B, _, T = s.shape
        y = F.conv1d(s.double(), conv.weight.detach().double(), conv.bias.detach().double(), padding=conv.padding[0])
        Tq, Tp = T // out.split, T // out.split + GAP
        for ph in range(out.split):
            for c in range(B + 1):
                out.store[ph][:, c * Tp: c * Tp + GAP] = 0
            for c in range(B):
                out.store[ph][:, c * Tp + GAP: c * Tp + GAP + Tq] = y[c][:, ph::out.split]

    def to_planar(self, x, out):
        B, Cn, T = x.shape
        Tp = T + GAP
        for c in range(B + 1):
            out.store[0][:, c * Tp: c * Tp + GAP] = 0
        for c in range(B):
            out.store[0][:, c * Tp + GAP: c * Tp + GAP + T] = x[c].double()

    def from_planar(self, x, Tout):
        Tp = x.T + GAP
        return torch.stack([x.store[0][:, c * Tp + GAP: c * Tp + GAP + Tout] for c in range(x.B)]).float()

    def tail8(self, x, rb, final, T):
        with torch.no_grad():
            h = self.from_planar(x, x.T)
            z = F.elu(rb.conv2(F.elu(rb.conv1(h))) + h)
            d = final(z)
        return d[:, :, :T] if d.shape[-1] >= T else F.pad(d, (0, T - d.shape[-1]))

    def run(self, g, srcs, B, T, elu, residual, mode, out, out_split=1, ct=None, out_T=0, cout=0, g2=None, skip=None):
        if g2 is not None:
            # contract of the fused residual block = the two calls it replaces
            assert mode == PC.OUT_PLANAR and PC.fusable(g, g2) and elu
            u = EmuPlanar(g.n_total, B, T)
            self.run(g, srcs, B, T, True, None, PC.OUT_PLANAR, u)
            self.calls[-1] = self.calls[-1] + ("fused",)
            s2 = [(u, 0)] + ([skip] if skip is not None else [])
            assert len(s2) == len(g2.srcs)
            return self.run(g2, s2, B, T, True, residual, PC.OUT_PLANAR, out, out_split=out_split)
        self.calls.append((mode, g.n_total, g.nc, [s[1:] for s in g.srcs]))
        Tp = T + GAP
        R = B * Tp + GAP
        nc, nch = g.nc, g.n_total // g.nc
        assert len(g.chunk_off) == nch and g.wd.shape[0] == nch and g.wd.shape[3] == nc
        acc = g.bias.double().repeat(R, 1)
        wd = g.wd.double()
        for j in range(nch):
            sl = 0
            for i, (sid, cin, off, taps) in enumerate(g.srcs):
                buf, ph = srcs[i]
                assert (buf.C, buf.B, buf.T) == (cin, B, T)
                assert off >= -GAP and off + taps - 1 <= GAP
                A = torch.cat([torch.full((cin, FRONT), float("nan"), dtype=torch.float64), buf.store[ph]], dim=1)
                o = off + (g.chunk_off[j] if i == 0 else 0)
                for kc in range(cin // 16):
                    for tp in range(taps):
                        rows = A[kc * 16:(kc + 1) * 16, FRONT + o + tp: FRONT + o + tp + R]
                        acc[:, j * nc:(j + 1) * nc] += rows.T @ wd[j, sl]
                        sl += 1
            assert sl == wd.shape[1]
        m = torch.arange(R)
        c, r = m // Tp, m % Tp
        t = r - GAP
        real = (c < B) & (t >= 0)
        if residual is not None:
            assert (residual.C, residual.B, residual.T) == (g.n_total, B, T)
            acc[real] += residual.store[0][:, :R].T[real]
        if elu:
            acc = torch.where(acc > 0, acc, torch.expm1(acc))
        if mode == PC.OUT_PLANAR:
            assert out.C == g.n_total and out.B == B
            if out_split <= 1:
                assert out.T == T
                out.store[0][:, :R] = torch.where(real[:, None], acc, torch.zeros_like(acc)).T
            else:
                sp = out_split
                assert T % sp == 0 and out.T == T // sp and out.split == sp
                oTp = T // sp + GAP
                used = [p for p in range(sp) if p <= 1 or p == sp - 1]
                for p in used:
                    sel = real & (t % sp == p)
                    out.store[p][:, (c * oTp + GAP + t // sp)[sel]] = acc[sel].T
                    gsel = ~real
                    out.store[p][:, (c * oTp + r)[gsel]] = 0
        elif mode == PC.OUT_CONVT:
            s, p, co_n = ct
            assert g.n_total == s * co_n and out.C == co_n and out.T == out_T and out.B == B
            oTp = out_T + GAP
            layout = g.cols if g.cols is not None else PC.convt_columns(s, co_n, 1)
            assert [PC.convt_decode(n, co_n, g.interleave) for n in range(g.n_total)] == layout   # the kernel's decoding
            ph_of = torch.tensor([c_[0] for c_ in layout])
            co_of = torch.tensor([c_[1] for c_ in layout])
            for ph in range(s):
                sel_cols = (ph_of == ph).nonzero().flatten()
                assert torch.equal(co_of[sel_cols].sort().values, torch.arange(co_n))
                cols = torch.empty(R, co_n, dtype=acc.dtype)
                cols[:, co_of[sel_cols]] = acc[:, sel_cols]
                tout = s * t + ph
                v = (tout >= 0) & (c < B) & (tout < out_T)
                out.store[0][:, (c * oTp + GAP + tout)[v]] = cols[v].T
                z = (tout < 0) & (tout >= -GAP)
                out.store[0][:, (c * oTp + GAP + tout)[z]] = 0
                tout2 = s * T + ph
                if tout2 < out_T:
                    v2 = (r == 0) & (c >= 1)
                    out.store[0][:, ((c - 1) * oTp + GAP + tout2)[v2]] = cols[v2].T
        else:
            assert out.shape == (B, cout, out_T) and out_T <= T
            sel = real & (t < out_T)
            out[c[sel], :, t[sel]] = acc[sel][:, :cout]
