"""Precision probe for the split-fp16 tensor-core evaluation of the distance MLP (run on the GPU box).

Compares, against an fp64 evaluation of the shipped network (forward + analytic input gradient):
  * fp32 : torch fp32 matmuls with TF32 off (FFMA / SGEMM arithmetic -- what exact_mlp.cu does)
  * split: every operand a = hi + lo * 2^-11 with hi = fp16(a), lo = fp16((a - hi) * 2^11); three fp16 tensor-core
           products with fp32 accumulation, D1 = hi.hi, D2 = hi.lo + lo.hi, result D1 + D2 * 2^-11
(cuBLAS runs the fp16 products on tcgen05 kind::f16 with the accumulator in TMEM, like csrc/tc_exact.cu).
Prints max / rms errors relative to the rms of the fp64 values."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden_util import load_weights  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
S = 2048.0


def split(a):
    hi = a.half()
    lo = ((a - hi.float()) * S).half()
    return hi, lo


def mm_split(a, wt):            # a (R, K) fp32, wt (N, K) fp32 -> a @ wt.T
    ah, al = split(a)
    wh, wl = split(wt)
    d1 = torch.mm(ah, wh.t(), out_dtype=torch.float32)
    d2 = torch.mm(ah, wl.t(), out_dtype=torch.float32) + torch.mm(al, wh.t(), out_dtype=torch.float32)
    return d1 + d2 / S


def mm_split_comp(a, wt):
    """mm_split with the accumulator-truncation compensation of tc_exact.cu: the tensor core truncates its fp32
    accumulator toward zero at every MMA (K = 16) step; over a chain of n steps that shrinks D1 by ~(1.65e-8 n + 2.1e-8)
    of its value on average (measured below), which the epilogue multiplies back."""
    ah, al = split(a)
    wh, wl = split(wt)
    d1 = torch.mm(ah, wh.t(), out_dtype=torch.float32)
    d2 = torch.mm(ah, wl.t(), out_dtype=torch.float32) + torch.mm(al, wh.t(), out_dtype=torch.float32)
    steps = (a.shape[1] + 15) // 16
    c = 1.65e-8 * steps + 2.1e-8
    return torch.addcmul(d1 + d2 / S, d1, torch.tensor(c, device=a.device, dtype=torch.float32))


def mm_fp32(a, wt):
    return a @ wt.t()


def mm_fp64(a, wt):
    return a.double() @ wt.double().t()


def net_eval(W, b, x, mm, dt):
    """x (R, nin) -> z (R, O), grad of z[argmin] w.r.t. x (R, nin); mm does every contraction."""
    nin = x.shape[1]
    xe = x.to(dt)
    enc = torch.cat([xe, torch.sin(xe), torch.cos(xe)], 1)
    h = enc if dt == torch.float64 else enc.float()
    masks = []
    for l in range(4):
        h = mm(h, W[l]).to(dt) + b[l].to(dt)
        masks.append(h > 0)
        h = torch.relu(h)
    z = mm(h, W[4]).to(dt) + b[4].to(dt)
    return z, masks, enc


def grad_eval(W, z_idx, masks, enc, x, mm, dt):
    nin = x.shape[1]
    g = W[4].to(dt)[z_idx] * masks[3]
    for l in (3, 2, 1):
        g = mm(g if dt == torch.float64 else g.float(), W[l].t().contiguous()).to(dt) * masks[l - 1]
    a = mm(g if dt == torch.float64 else g.float(), W[0].t().contiguous()).to(dt)
    xe = x.to(dt)
    return a[:, :nin] + torch.cos(xe) * a[:, nin:2 * nin] - torch.sin(xe) * a[:, 2 * nin:]


def stats(name, v, ref):
    e = (v.double() - ref).abs()
    scale = ref.pow(2).mean().sqrt()
    rel = e / ref.abs().clamp_min(1e-3 * scale)
    print(f"  {name:28s} max|err|/rms {e.max().item() / scale:9.2e}   rms|err|/rms {e.pow(2).mean().sqrt().item() / scale:9.2e}"
          f"   max rel {rel.max().item():9.2e}   frac(rel<1e-5) {(rel < 1e-5).double().mean().item():.5f}")


def main():
    torch.manual_seed(0)
    # ---- raw GEMM, K = 256
    a = torch.relu(torch.randn(8192, 256, device=dev))
    w = torch.randn(256, 256, device=dev) * 0.08
    ref = mm_fp64(a, w)
    print("GEMM 8192x256x256 (relu-normal activations, N(0, 0.08) weights)")
    stats("fp32 (FFMA)", mm_fp32(a, w), ref)
    stats("split fp16 x3", mm_split(a, w), ref)
    ah, al = split(a)
    wh, wl = split(w)
    stats("split + compensation", mm_split_comp(a, w), ref)
    stats("hi.hi only (plain fp16)", torch.mm(ah, wh.t(), out_dtype=torch.float32), ref)
    # is the extra error of the split path a bias of the tensor-core accumulator (truncation), and does a shorter
    # in-accumulator chain (K split in chunks, partial sums added in fp32 RN outside) remove it?
    def mm_split_chunks(a, wt, chunks):
        out = None
        step = a.shape[1] // chunks
        for c0 in range(0, a.shape[1], step):
            part = mm_split(a[:, c0:c0 + step].contiguous(), wt[:, c0:c0 + step].contiguous())
            out = part if out is None else out + part
        return out
    for chunks in (1, 2, 4, 16):
        v = mm_split_chunks(a, w, chunks)
        e = (v.double() - ref)
        big = ref.abs() > ref.pow(2).mean().sqrt()
        print(f"  split, K in {chunks:2d} chunks: rms|err|/rms {e.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item():9.2e}"
              f"   mean(err*sign(ref))/rms {(e * ref.sign()).mean().item() / ref.pow(2).mean().sqrt().item():10.2e}"
              f"   mean rel err on |ref|>rms {(e[big] / ref[big]).mean().item():10.2e}")
    e = (mm_fp32(a, w).double() - ref)
    print(f"  fp32 FFMA             : mean(err*sign(ref))/rms {(e * ref.sign()).mean().item() / ref.pow(2).mean().sqrt().item():10.2e}")
    for net, d in (("franka", 7), ("planar7", 7), ("planar2", 2)):
        W, b = load_weights(net)
        W = [t.to(dev) for t in W]
        b = [t.to(dev) for t in b]
        nin = W[0].shape[1] // 3
        R = 65536
        x = torch.empty(R, nin, device=dev)
        x[:, :d] = (torch.rand(R, d, device=dev) * 2 - 1) * 2.8
        x[:, d:] = (torch.rand(R, nin - d, device=dev) * 2 - 1) * (1.0 if net == "franka" else 6.0)
        z64, m64, e64 = net_eval(W, b, x, mm_fp64, torch.float64)
        idx = z64.argmin(1)
        g64 = grad_eval(W, idx, m64, e64, x, mm_fp64, torch.float64)
        print(f"{net}: {R} rows, |z| rms {z64.pow(2).mean().sqrt().item():.3f}, |grad| rms {g64.pow(2).mean().sqrt().item():.3f}")
        acts = [e64.abs().max().item()]
        h = e64
        for l in range(4):
            h = torch.relu(h @ W[l].double().t() + b[l].double())
            acts.append(h.abs().max().item())
        print("  max |activation| per layer:", " ".join(f"{v:.3g}" for v in acts),
              " max|W|:", " ".join(f"{t.abs().max().item():.3g}" for t in W))
        for name, mm in (("fp32 (FFMA)", mm_fp32), ("split fp16 x3", mm_split), ("split + comp", mm_split_comp)):
            z, m, e = net_eval(W, b, x, mm, torch.float32)
            stats(name + " z", z, z64)
            g = grad_eval(W, idx, m, e, x, mm, torch.float32)
            # rows whose ReLU pattern differs from fp64 sit on a kink: report them separately
            same = torch.ones(R, dtype=torch.bool, device=dev)
            for l in range(4):
                same &= (m[l] == m64[l]).all(1)
            stats(name + " grad (same relu)", g[same], g64[same])
            print(f"    rows with a flipped ReLU vs fp64: {(~same).sum().item()}   max |g| seen {g.abs().max().item():.3g}")


if __name__ == "__main__":
    main()
