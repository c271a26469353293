#!/usr/bin/env python
"""GPU probe of the TMA-fed tensor-core kernels (pf_conv_tma.cu), written to be diagnosable from one run without a
local GPU.  Groups (each one a separate process under `timeout`, so a trap in one does not take the rest down):

  onehot   forward conv whose kernel is a one-hot tap/channel selector: the output must be a shifted copy of the input;
           on mismatch the script searches which input pixel each wrong output row actually holds and prints the
           mapping it found next to the one it expected (pins the im2col tensor-map conventions);
  fwd      random convs, TMA kernels vs float64 and vs the cp.async kernels (same split-bf16 planes);
  dgrad    unit-stride dgrad likewise;
  wgrad    weight gradient likewise;
  levels   integer-level operands: forward with level x level (1 MMA) and level x split (2 MMAs), wgrad with levels.

usage: python tools/tma_probe.py <group> [...]; prints one line per case, `PROBE <group> ok|FAIL`."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pocketflow_b200 import ops  # noqa: E402

DEV = torch.device('cuda:0')
# n, h, w, c, k, r, s, stride, pad0 (top/left), pad1 (bottom/right)
CASES = [
    (2, 8, 8, 64, 64, 1, 1, 1, 0, 0),
    (2, 9, 7, 64, 64, 3, 3, 1, 1, 1),
    (3, 14, 14, 64, 128, 3, 3, 1, 1, 1),
    (2, 12, 12, 128, 128, 3, 3, 2, 0, 1),
    (2, 13, 11, 64, 64, 3, 3, 2, 1, 1),
    (2, 14, 14, 256, 512, 1, 1, 2, 0, 0),
    (1, 7, 7, 512, 512, 3, 3, 1, 1, 1),
    (2, 7, 7, 512, 2048, 1, 1, 1, 0, 0),
    (8, 56, 56, 64, 256, 1, 1, 1, 0, 0),
    (8, 56, 56, 256, 64, 1, 1, 1, 0, 0),
    (2, 14, 14, 64, 192, 3, 3, 1, 1, 1),
    (5, 10, 10, 64, 64, 5, 5, 1, 0, 0),
    (4, 28, 28, 128, 128, 3, 3, 1, 1, 1),
    # ResNet-50 shapes at batch 2 (the parity tests' configuration)
    (2, 56, 56, 64, 64, 3, 3, 1, 1, 1),
    (2, 56, 56, 128, 128, 3, 3, 2, 1, 1),
    (2, 28, 28, 128, 128, 3, 3, 1, 1, 1),
    (2, 14, 14, 256, 256, 3, 3, 1, 1, 1),
]


def geom(case):
    n, h, w, c, k, r, s, st, p0, p1 = case
    p = (h + p0 + p1 - r) // st + 1
    q = (w + p0 + p1 - s) // st + 1
    return p, q, ops.conv_desc(n, h, w, c, k, r, s, p, q, st, st, p0, p0)


def ref_conv(x, wt, case):
    n, h, w, c, k, r, s, st, p0, p1 = case
    xd = x.double().permute(0, 3, 1, 2)
    wd = wt.double().permute(3, 2, 0, 1)
    return F.conv2d(F.pad(xd, (p0, p1, p0, p1)), wd, stride=st).permute(0, 2, 3, 1)


def planes_of(t):
    pl = ops.Planes(t.numel(), DEV)
    ops.split_bf16(t.contiguous(), pl)
    return pl


def relerr(got, ref):
    return (got.double().cpu() - ref.cpu()).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def describe_bad(got, ref, tol, shape_names):
    err = (got.double().cpu() - ref.cpu()).abs()
    bad = err > tol * ref.abs().max()
    idx = bad.nonzero()
    out = ['bad %d of %d' % (int(bad.sum()), bad.numel())]
    for j, nm in enumerate(shape_names):
        vals, cnt = torch.unique(idx[:, j], return_counts=True)
        out.append('%s: %s' % (nm, ', '.join('%d(%d)' % (int(v), int(c)) for v, c in list(zip(vals, cnt))[:12])))
    for row in idx[:6]:
        t = tuple(int(v) for v in row)
        out.append('  at %s got %.6g ref %.6g' % (t, float(got[t]), float(ref[t])))
    return '\n    '.join(out)


def group_onehot():
    ok = True
    for case in CASES[:6]:
        n, h, w, c, k, r, s, st, p0, p1 = case
        p, q, d = geom(case)
        if not ops.conv2d_tc_tma_supported(d, 0):
            print('onehot', case, 'not TMA-eligible (skipped)')
            continue
        g = torch.Generator().manual_seed(sum(case))
        x = torch.randn(n, h, w, c, generator=g)
        X = x.to(DEV)
        xp = planes_of(X)
        for (tr, tq) in sorted({(0, 0), (r - 1, s - 1), (r // 2, s // 2), (0, s - 1)}):
            wt = torch.zeros(r, s, c, k)
            for co in range(k):
                wt[tr, tq, co % c, co] = 1.0
            tw = ops.TcWeights(d, DEV)
            tw.prepare(wt.to(DEV).contiguous())
            Y = torch.full((n, p, q, k), float('nan'), device=DEV)
            ops.conv2d_tc_set_feed(1)
            ops.conv2d_tc_fwd_planes(d, xp, tw, None, False, Y)
            torch.cuda.synchronize()
            ref = ref_conv(x, wt, case)
            e = relerr(Y, ref)
            good = e <= 1e-5
            ok &= good
            print('onehot %s tap (%d,%d): err %.3e %s' % (case, tr, tq, e, 'ok' if good else 'FAIL'))
            if not good:
                # which input pixel does each output position hold?  compare on the first min(c, k) channels
                Yc = Y.cpu()
                cc = min(c, k)
                xin = x[..., :cc].reshape(-1, cc)
                shown = 0
                for ni in range(n):
                    for yy in range(p):
                        for xx in range(q):
                            row = Yc[ni, yy, xx, :cc]
                            exp = ref[ni, yy, xx, :cc].float()
                            if torch.allclose(row, exp, atol=1e-4):
                                continue
                            if shown >= 24:
                                continue
                            shown += 1
                            if torch.isnan(row).any():
                                found = 'NaN (never written)'
                            elif row.abs().max() == 0:
                                found = 'zeros'
                            else:
                                dist = (xin - row[None]).abs().max(dim=1).values
                                j = int(dist.argmin())
                                found = 'x[n=%d,h=%d,w=%d] (d=%.2e)' % (j // (h * w), (j // w) % h, j % w, float(dist[j]))
                            ih, iw = yy * st - p0 + tr, xx * st - p0 + tq
                            want = 'x[n=%d,h=%d,w=%d]' % (ni, ih, iw) if 0 <= ih < h and 0 <= iw < w else 'zeros (padding)'
                            print('    out[n=%d,y=%d,x=%d] holds %s, expected %s' % (ni, yy, xx, found, want))
    return ok


def run_pair(fn):
    """fn(feed) -> tensor; returns (tma result, lsu result)"""
    ops.conv2d_tc_set_feed(1)
    a = fn()
    torch.cuda.synchronize()
    ops.conv2d_tc_set_feed(0)
    b = fn()
    torch.cuda.synchronize()
    ops.conv2d_tc_set_feed(1)
    return a, b


def group_fwd():
    ok = True
    for case in CASES:
        n, h, w, c, k, r, s, st, p0, p1 = case
        p, q, d = geom(case)
        if not ops.conv2d_tc_tma_supported(d, 0):
            print('fwd', case, 'not TMA-eligible (skipped)')
            continue
        g = torch.Generator().manual_seed(sum(case))
        x = torch.randn(n, h, w, c, generator=g)
        wt = torch.randn(r, s, c, k, generator=g) * (2.0 / (r * s * c)) ** 0.5
        bias, res = torch.randn(k, generator=g), torch.randn(n, p, q, k, generator=g)
        ref = torch.relu(ref_conv(x, wt, case) + bias.double()) + res.double()
        xp = planes_of(x.to(DEV))
        tw = ops.TcWeights(d, DEV)
        tw.prepare(wt.to(DEV).contiguous())
        B, R = bias.to(DEV), res.to(DEV)

        def f():
            Y = torch.full((n, p, q, k), float('nan'), device=DEV)
            ops.conv2d_tc_fwd_planes(d, xp, tw, B, True, Y, R)
            return Y
        a, b = run_pair(f)
        e, same = relerr(a, ref), bool(torch.equal(a, b))
        good = e <= 2e-5
        ok &= good
        print('fwd %s: err %.3e, bit-equal to cp.async kernel: %s %s' % (case, e, same, 'ok' if good else 'FAIL'))
        if not good:
            print('    ' + describe_bad(a, ref, 2e-5, ['n', 'y', 'x', 'k']))
    return ok


def group_dgrad():
    ok = True
    for case in CASES:
        n, h, w, c, k, r, s, st, p0, p1 = case
        p, q, d = geom(case)
        if not ops.conv2d_tc_tma_supported(d, 1):
            print('dgrad', case, 'not TMA-eligible (skipped)')
            continue
        g = torch.Generator().manual_seed(sum(case) + 7)
        wt = torch.randn(r, s, c, k, generator=g) * (2.0 / (r * s * c)) ** 0.5
        dy = torch.randn(n, p, q, k, generator=g)
        xd = torch.zeros(n, c, h, w, dtype=torch.float64, requires_grad=True)
        yd = F.conv2d(F.pad(xd, (p0, p1, p0, p1)), wt.double().permute(3, 2, 0, 1), stride=st)
        yd.backward(dy.double().permute(0, 3, 1, 2))
        ref = xd.grad.permute(0, 2, 3, 1)
        dyp = planes_of(dy.to(DEV))
        tw = ops.TcWeights(d, DEV)
        tw.prepare(wt.to(DEV).contiguous())

        def f():
            DX = torch.full((n, h, w, c), 1.0, device=DEV)
            ops.conv2d_tc_dgrad_planes(d, dyp, tw, False, DX)
            ops.conv2d_tc_dgrad_planes(d, dyp, tw, True, DX)
            return DX
        a, b = run_pair(f)
        e, same = relerr(a, 2 * ref), bool(torch.equal(a, b))
        good = e <= 2e-5
        ok &= good
        print('dgrad %s: err %.3e, bit-equal to cp.async kernel: %s %s' % (case, e, same, 'ok' if good else 'FAIL'))
        if not good:
            print('    ' + describe_bad(a, 2 * ref, 2e-5, ['n', 'h', 'w', 'c']))
    return ok


def wgrad_ref(x, dy, case):
    n, h, w, c, k, r, s, st, p0, p1 = case
    wd = torch.zeros(k, c, r, s, dtype=torch.float64, requires_grad=True)
    yd = F.conv2d(F.pad(x.double().permute(0, 3, 1, 2), (p0, p1, p0, p1)), wd, stride=st)
    yd.backward(dy.double().permute(0, 3, 1, 2))
    return wd.grad.permute(2, 3, 1, 0)


def group_wgrad():
    ok = True
    for case in CASES:
        n, h, w, c, k, r, s, st, p0, p1 = case
        p, q, d = geom(case)
        if not ops.conv2d_tc_tma_supported(d, 2):
            print('wgrad', case, 'not TMA-eligible (skipped)')
            continue
        g = torch.Generator().manual_seed(sum(case) + 1)
        x = torch.randn(n, h, w, c, generator=g)
        dy = torch.randn(n, p, q, k, generator=g)
        ref = wgrad_ref(x, dy, case)
        xp, dyp = planes_of(x.to(DEV)), planes_of(dy.to(DEV))
        ws = torch.empty(max(ops.conv2d_tc_wgrad_planes_workspace_floats(d), 4), device=DEV)

        def f():
            DW = torch.full((r, s, c, k), 5.0, device=DEV)
            ops.conv2d_tc_wgrad_planes(d, xp, dyp, ws, DW)
            return DW
        a, b = run_pair(f)
        e, same = relerr(a, ref), bool(torch.equal(a, b))
        good = e <= 2e-5
        ok &= good
        print('wgrad %s: err %.3e, bit-equal to cp.async kernel: %s %s' % (case, e, same, 'ok' if good else 'FAIL'))
        if not good:
            print('    ' + describe_bad(a, ref, 2e-5, ['r', 's', 'c', 'k']))
    return ok


def channel_sums(levels, seg=128):
    """[n,h,w,c] -> [n*h*w, ceil(c/seg)] sums over channel segments (what the producer kernel emits)"""
    n, h, w, c = levels.shape
    nseg = (c + seg - 1) // seg
    v = levels.reshape(-1, c).double()
    out = torch.zeros(v.shape[0], nseg, dtype=torch.float64)
    for i in range(nseg):
        out[:, i] = v[:, i * seg:(i + 1) * seg].sum(1)
    return out.float().contiguous(), nseg


def group_levels():
    ok = True
    ops.conv2d_tc_set_feed(1)
    for case in CASES:
        n, h, w, c, k, r, s, st, p0, p1 = case
        p, q, d = geom(case)
        if not (ops.conv2d_tc_tma_supported(d, 0) and ops.conv2d_tc_tma_supported(d, 2)):
            print('levels', case, 'not TMA-eligible (skipped)')
            continue
        g = torch.Generator().manual_seed(sum(case) + 3)
        for bits, per_channel in ((8, True), (4, False)):
            kq = (1 << bits) - 1
            centre = float(1 << (bits - 1))
            j = torch.randint(0, 256, (n, h, w, c), generator=g).float()          # activation levels
            j[torch.rand(n, h, w, c, generator=g) < 0.4] = 0.0                     # ReLU-like: many exact zeros
            s_a = 0.0173
            lv = torch.randint(0, kq + 1, (r, s, c, k), generator=g).float()       # weight levels
            nb = k if per_channel else 1
            alpha = (torch.rand(nb, generator=g) * 0.5 + 0.05)
            beta = -alpha * (0.3 + 0.4 * torch.rand(nb, generator=g))
            rk = np.float32(1.0) / np.float32(kq)
            qw = (alpha.double() * float(rk)) * lv.double() + beta.double()          # broadcast over the last axis
            qa = j.double() * s_a
            bias, res = torch.randn(k, generator=g), torch.randn(n, p, q, k, generator=g)
            ref = torch.relu(ref_conv(qa, qw, case) + bias.double()) + res.double()
            # operands
            hdr = torch.from_numpy(np.array([(s_a, 1)], dtype=ops.ACT_HDR).view(np.uint8)).to(DEV)
            csum, nseg = channel_sums(j)
            csum = csum.to(DEV)
            apl = ops.Planes(j.numel(), DEV)
            apl.hi.copy_(j.reshape(-1).to(torch.bfloat16))
            apl.lo.fill_(float('nan'))                                               # must never be read
            act = ops.tc_act(apl, hdr, csum, nseg)
            wl = (lv - centre).permute(3, 0, 1, 2).reshape(k, r * s * c).to(torch.bfloat16).contiguous().to(DEV)
            A, Bt = alpha.to(DEV), beta.to(DEV)
            if A.numel() % 4:
                A = torch.cat([A, torch.zeros(4 - A.numel() % 4, device=DEV)])
                Bt = torch.cat([Bt, torch.zeros(4 - Bt.numel() % 4, device=DEV)])
            wt_ = ops.tc_wt(wl, None, A, Bt, per_channel, bits)
            Y = torch.full((n, p, q, k), float('nan'), device=DEV)
            ops.conv2d_tc_fwd_ex(d, act, wt_, bias.to(DEV), True, Y, res.to(DEV))
            torch.cuda.synchronize()
            e = relerr(Y, ref)
            good = e <= 1e-5
            ok &= good
            print('levels fwd 1-MMA %s W%d %s: err %.3e %s' % (case, bits, 'per-channel' if per_channel else 'per-layer', e,
                                                               'ok' if good else 'FAIL'))
            if not good:
                print('    ' + describe_bad(Y, ref, 1e-5, ['n', 'y', 'x', 'k']))
        # level activations x split-bf16 weights (2 MMAs), scalar scale in the epilogue
        wt = torch.randn(r, s, c, k, generator=g) * (2.0 / (r * s * c)) ** 0.5
        tw = ops.TcWeights(d, DEV)
        tw.prepare(wt.to(DEV).contiguous())
        ref2 = ref_conv(qa, wt, case)
        Y = torch.full((n, p, q, k), float('nan'), device=DEV)
        ops.conv2d_tc_fwd_ex(d, act, ops.tc_wt(tw.f_hi, tw.f_lo), None, False, Y)
        torch.cuda.synchronize()
        e = relerr(Y, ref2)
        good = e <= 2e-5
        ok &= good
        print('levels fwd 2-MMA %s: err %.3e %s' % (case, e, 'ok' if good else 'FAIL'))
        # header says two planes: the same operand as hi / lo with scale 1 must give the plain result
        hdr2 = torch.from_numpy(np.array([(1.0, 2)], dtype=ops.ACT_HDR).view(np.uint8)).to(DEV)
        qpl = planes_of(qa.float().to(DEV))
        Y2 = torch.full((n, p, q, k), float('nan'), device=DEV)
        ops.conv2d_tc_fwd_ex(d, ops.tc_act(qpl, hdr2), ops.tc_wt(tw.f_hi, tw.f_lo), None, False, Y2)
        torch.cuda.synchronize()
        e = relerr(Y2, ref2)
        good = e <= 2e-5
        ok &= good
        print('levels fwd hdr=2 planes %s: err %.3e %s' % (case, e, 'ok' if good else 'FAIL'))
        # wgrad: levels (x) split dy, scaled by s_a
        dy = torch.randn(n, p, q, k, generator=g)
        refw = wgrad_ref(qa, dy, case)
        dyp = planes_of(dy.to(DEV))
        ws = torch.empty(max(ops.conv2d_tc_wgrad_planes_workspace_floats(d), 4), device=DEV)
        DW = torch.full((r, s, c, k), 5.0, device=DEV)
        ops.conv2d_tc_wgrad_ex(d, act, ops.tc_act(dyp), ws, DW)
        torch.cuda.synchronize()
        e = relerr(DW, refw)
        good = e <= 2e-5
        ok &= good
        print('levels wgrad 2-MMA %s: err %.3e %s' % (case, e, 'ok' if good else 'FAIL'))
        if not good:
            print('    ' + describe_bad(DW, refw, 2e-5, ['r', 's', 'c', 'k']))
    return ok


GROUPS = dict(onehot=group_onehot, fwd=group_fwd, dgrad=group_dgrad, wgrad=group_wgrad, levels=group_levels)

if __name__ == '__main__':
    rc = 0
    for name in sys.argv[1:]:
        try:
            good = GROUPS[name]()
        except Exception as e:   # noqa: BLE001 — report and keep the exit code
            print('PROBE %s EXCEPTION %s: %s' % (name, type(e).__name__, e))
            good = False
        print('PROBE %s %s' % (name, 'ok' if good else 'FAIL'))
        rc |= 0 if good else 1
    sys.exit(rc)
