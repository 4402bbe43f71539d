"""Diagnostics of the tcgen05 gather-GEMM (csrc/sparse_conv.cu: k_gather_gemm_tc) on structured inputs: which operand
layout is wrong shows in WHERE the output differs, not only that it does.  Run on the GPU box; prints one JSON line per
case and writes gpurun_out/<tag>_tc_diag.json.   usage: python tools/tc_diag.py [tag]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nksr_b200.unet import gather_gemm, round_tf32  # noqa: E402


def tc(x, idx, w, bias=None, res=None, relu=False):
    return gather_gemm(x, idx, round_tf32(w).transpose(1, 2).contiguous(), bias, res, relu, tf32=3)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "dev"
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    out = []

    def record(name, got, ref, extra=None):
        err = float((got - ref).abs().max())
        scale = float(ref.abs().max()) + 1e-30
        bad = (got - ref).abs() > 4e-3 * scale
        rec = dict(case=name, max_err=err, scale=scale, rel=err / scale, bad_frac=float(bad.float().mean()),
                   finite=bool(torch.isfinite(got).all()))
        if bad.any():
            rows = torch.nonzero(bad.any(dim=1)).squeeze(1)
            cols = torch.nonzero(bad.any(dim=0)).squeeze(1)
            rec.update(bad_rows=rows[:16].tolist(), n_bad_rows=int(rows.numel()), bad_cols=cols[:64].tolist())
        if extra:
            rec.update(extra)
        out.append(rec)
        print(json.dumps(rec), flush=True)

    for c_in, c_out in [(32, 32), (32, 64), (64, 64), (128, 64), (32, 96)]:
        n = 300
        x = torch.randn((n, c_in), generator=g).to(dev)
        ident = torch.arange(n, dtype=torch.int32, device=dev)[:, None].contiguous()
        # 1. one tap, identity gather, W = [I | 0] or stacked identities: y[:, j] = x[:, j % c_in] -- A, D layouts
        w = torch.zeros((1, c_in, c_out))
        for j in range(c_out):
            w[0, j % c_in, j] = 1.0
        w = w.to(dev)
        ref = round_tf32(x)[:, [j % c_in for j in range(c_out)]]
        got = tc(x, ident, w)
        extra = None
        if c_in == c_out == 32 and not torch.allclose(got, ref, atol=2e-3):
            # where did column j of x go?  (row 5, a generic row)
            m = (got[5][None, :] - x[5][:, None]).abs() < 2e-3
            extra = dict(col_map_row5=[torch.nonzero(m[j]).squeeze(1).tolist() for j in range(32)])
        record(f"identity_{c_in}x{c_out}", got, ref, extra)
        # 2. one-hot rows: y[i] = W[0][i % c_in, :] -- B layout
        x1 = torch.zeros((n, c_in))
        x1[torch.arange(n), torch.arange(n) % c_in] = 1.0
        w2 = (torch.randn((1, c_in, c_out), generator=g)).to(dev)
        got = tc(x1.to(dev), ident, w2)
        record(f"onehot_{c_in}x{c_out}", got, round_tf32(w2)[0][torch.arange(n) % c_in])
        # 3. random, 27 taps with holes
        idx = torch.randint(-1, n, (1000, 27), generator=g).to(torch.int32).to(dev)
        w3 = (torch.randn((27, c_in, c_out), generator=g) / (27 * c_in) ** 0.5).to(dev)
        b = torch.randn(c_out, generator=g).to(dev)
        res = torch.randn((1000, c_out), generator=g).to(dev)
        ref = gather_gemm(x, idx, w3, b, res, True, impl="torch")
        record(f"random27_{c_in}x{c_out}", tc(x, idx, w3, b, res, True), ref)
    # timing of one level-0-sized convolution against the mma.sync kernel
    n = 2_000_000
    x = torch.randn((n, 32), device=dev)
    nb = (torch.arange(n, device=dev)[:, None] + torch.arange(-13, 14, device=dev)[None, :] * 37)
    idx = torch.where((nb >= 0) & (nb < n), nb, torch.full_like(nb, -1)).to(torch.int32).contiguous()
    w = torch.randn((27, 32, 32), device=dev) / 30
    wt, wr = round_tf32(w).transpose(1, 2).contiguous(), round_tf32(w)
    for name, fn in [("mma_sync_tf32", lambda: gather_gemm(x, idx, wr, None, None, True, tf32=2)),
                     ("tcgen05_tf32", lambda: gather_gemm(x, idx, wt, None, None, True, tf32=3)),
                     ("ffma_fp32", lambda: gather_gemm(x, idx, w, None, None, True, tf32=0))]:
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rec = dict(case=f"time_{name}", n_out=n, c=32, ms=ms, tflops=2 * n * 27 * 32 * 32 / ms / 1e9)
        out.append(rec)
        print(json.dumps(rec), flush=True)
    a, b_ = gather_gemm(x, idx, wr, None, None, True, tf32=2), gather_gemm(x, idx, wt, None, None, True, tf32=3)
    record("big_tc_vs_mma_sync", b_, a)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/{tag}_tc_diag.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
