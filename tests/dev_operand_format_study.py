"""(dev tool, CPU only) Which tensor-core operand format keeps the PPO gradients of the 64x64 network inside north_star's 1e-4?

The update kernel (csrc/update_tc.cu) runs four GEMM groups per 128-sample tile on tcgen05: forward z2 = h1 W2^T, dgrad dh1 = dz2 W2,
wgrad dW2 = dz2^T h1 and the tail (dW1 = dz1^T x, dW_heads = g^T h2); everything else (layer 1, heads, softmax, loss, activation
derivatives) is fp32 SIMT.  Today every product is a 3xTF32 split (kind::tf32, K = 8 per instruction).  DESIGN.md §7 item 1 proposes a
bf16 split (kind::f16, K = 16 per instruction: half the MMA instructions and half the operand bytes per product term).  This script
restates the network's forward / backward by hand, runs ONLY those four GEMMs through an emulated operand format (operands rounded /
split exactly as the hardware would see them, products and sums in fp32 like the TMEM accumulator), and compares the gradients with an
fp64 autograd reference (oracle/policy.py) on the same minibatch:

  fp32        operands untouched (what the fp32 FMA-pipe kernel computes)
  tf32x3      hi = rn_tf32(x) (tc_common.cuh::tf32_rn), lo = x - hi, read by the tensor core as its top 19 bits;  hi*hi + lo*hi + hi*lo   3 MMAs, K=8
  tf32x1      the fp32 operand as the tensor core reads it (top 19 bits: TRUNCATED)                                          1 MMA,  K=8
  bf16x3      b0 = rn_bf16(x), b1 = rn_bf16(x - b0);  b0*b0 + b1*b0 + b0*b1                                                   3 MMAs, K=16
  bf16x6      b0, b1, b2 = three bf16 terms;  all products down to 2^-24: b0b0 + b0b1 + b1b0 + b1b1 + b0b2 + b2b0            6 MMAs, K=16
  bf16x1      b0*b0                                                                                                           1 MMA,  K=16

    python tests/dev_operand_format_study.py [--n 262144]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import policy as P


def trunc19(x: torch.Tensor) -> torch.Tensor:
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def rn_tf32(x: torch.Tensor) -> torch.Tensor:
    """tc_common.cuh::tf32_rn: round-half-away on the 13 dropped mantissa bits."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def split(x, fmt):
    if fmt == "tf32x1":
        return [trunc19(x)]                        # a bare fp32 operand: the tensor core reads its top 19 bits
    if fmt.startswith("tf32"):
        hi = rn_tf32(x)                            # what update_tc.cu stores: hi rounded to nearest, lo = x - hi exact in fp32 ...
        return [hi, trunc19(x - hi)]               # ... of which the tensor core reads the top 19 bits
    b0 = bf16(x)
    b1 = bf16(x - b0)
    return [b0, b1, bf16(x - b0 - b1)]


TERMS = {"tf32x3": [(0, 0), (1, 0), (0, 1)], "tf32x1": [(0, 0)], "bf16x3": [(0, 0), (1, 0), (0, 1)], "bf16x1": [(0, 0)],
         "bf16x6": [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (0, 2)]}


def mm(a, b, fmt):
    """a @ b with both operands in the emulated format, fp32 products and accumulation."""
    if fmt == "fp32":
        return a @ b
    sa, sb = split(a, fmt), split(b, fmt)
    out = None
    for i, j in reversed(TERMS[fmt]):              # small terms first, like a compensated sum would (the order barely matters here)
        t = sa[i] @ sb[j]
        out = t if out is None else out + t
    return out


def ppo_grads(p, obs, actions, old_logp, values_old, adv, ret, fmt, clip=0.2, clip_vf=0.2, vf_coef=0.5, ent_coef=0.01):
    """Hand-written forward / backward of the PPO loss (oracle/policy.py::ppo_loss, batch-normalised advantages), GEMMs through mm()."""
    n = obs.shape[0]
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    z1 = obs @ p["w1"].t() + p["b1"]                       # SIMT fp32 in the kernel (K = 4)
    h1 = torch.relu(z1)
    z2 = mm(h1, p["w2"].t().contiguous(), fmt) + p["b2"]   # fwd group
    h2 = torch.relu(z2)
    logits = h2 @ p["wp"].t() + p["bp"]                    # heads: SIMT fp32
    v = (h2 @ p["wv"].t() + p["bv"]).squeeze(-1)
    logp_all = logits - logits.logsumexp(-1, keepdim=True)
    probs = logp_all.exp()
    logp = logp_all.gather(-1, actions.view(-1, 1)).squeeze(-1)
    ratio = (logp - old_logp).exp()
    s1, s2 = adv * ratio, adv * ratio.clamp(1 - clip, 1 + clip)
    # d(policy_loss)/d(logp): -mean(min(s1, s2)); the clamp passes no gradient where it is active
    use1 = s1 <= s2
    inside = (ratio >= 1 - clip) & (ratio <= 1 + clip)
    dlogp = -(torch.where(use1, adv * ratio, torch.where(inside, adv * ratio, torch.zeros_like(ratio)))) / n
    onehot = torch.zeros_like(probs).scatter_(1, actions.view(-1, 1), 1.0)
    dlogits = dlogp.unsqueeze(-1) * (onehot - probs)
    # entropy bonus: loss += ent_coef * (-H);  dH/dlogits = -p * (log p + H)
    H = -(probs * logp_all).sum(-1, keepdim=True)
    dlogits += ent_coef * (probs * (logp_all + H)) / n
    # clipped value loss: max((v - ret)^2, (v_clipped - ret)^2)
    vd = v - values_old
    vc = values_old + vd.clamp(-clip_vf, clip_vf)
    lu, lc = (v - ret) ** 2, (vc - ret) ** 2
    inside_v = (vd >= -clip_vf) & (vd <= clip_vf)
    dv = vf_coef * torch.where(lu >= lc, 2 * (v - ret), torch.where(inside_v, 2 * (vc - ret), torch.zeros_like(v))) / n
    g = torch.cat([dlogits, dv.unsqueeze(-1)], dim=1)                              # (n, A + 1): gradient at the head outputs
    wh = torch.cat([p["wp"], p["wv"]], dim=0)                                      # (A + 1, 64)
    dwh = mm(g.t().contiguous(), h2, fmt)                                          # tail group
    dh2 = g @ wh                                                                   # SIMT fp32 (K = 3)
    dz2 = dh2 * (z2 > 0)
    dw2 = mm(dz2.t().contiguous(), h1, fmt)                                        # wgrad group
    dh1 = mm(dz2, p["w2"], fmt)                                                    # dgrad group
    dz1 = dh1 * (z1 > 0)
    dw1 = mm(dz1.t().contiguous(), obs, fmt)                                       # tail group
    A = p["wp"].shape[0]
    grads = {"w1": dw1, "b1": dz1.sum(0), "w2": dw2, "b2": dz2.sum(0), "wp": dwh[:A], "bp": g[:, :A].sum(0), "wv": dwh[A:], "bv": g[:, A:].sum(0)}
    return torch.cat([grads[k].reshape(-1) for k in P.PARAM_ORDER])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=262144)
    ap.add_argument("--hidden", type=int, default=64, help="width of both hidden layers (64: the bench network; 128 / 256: the C4 presets)")
    ap.add_argument("--obs-dim", type=int, default=4)
    ap.add_argument("--actions", type=int, default=2)
    args = ap.parse_args()
    n, D, A, Hd = args.n, args.obs_dim, args.actions, args.hidden
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(7)
    p = P.random_params(D, (Hd, Hd), A, seed=3, has_value=True)
    obs = torch.randn(n, D, generator=g)
    actions = torch.randint(0, A, (n,), generator=g)
    with torch.no_grad():
        logits, v = P.forward(p, obs)
        lp = (logits - logits.logsumexp(-1, keepdim=True)).gather(-1, actions.view(-1, 1)).squeeze(-1)
    old_logp = lp + 0.2 * torch.randn(n, generator=g)
    values_old = v + 0.3 * torch.randn(n, generator=g)
    adv = torch.randn(n, generator=g) * 2 + 0.3
    ret = values_old + adv
    hp = dict(clip_range=0.2, clip_range_vf=0.2, vf_coef=0.5, ent_coef=0.01, normalize_adv=True)
    p64 = {k: t.double() for k, t in p.items()}
    _, ref, _ = P.loss_and_grads(P.ppo_loss, p64, obs.double(), actions, old_logp.double(), values_old.double(), adv.double(), ret.double(), **hp)
    ref = ref.numpy()
    scale = np.abs(ref).max()
    print(f"PPO gradients of the {Hd}x{Hd} network (obs {D}, actions {A}), {n:,}-sample minibatch, against an fp64 autograd reference (gradient scale max|g| = {scale:.3e};"
          f" north_star tolerance: 1e-4 of it)\n")
    print("| operand format of the four tensor-core GEMM groups | MMAs per product (K per instruction) | max abs error / scale | relative L2 error |")
    print("|---|---|---|---|")
    cost = {"fp32": "- (FMA pipe)", "tf32x3": "3 (K=8)", "tf32x1": "1 (K=8)", "bf16x3": "3 (K=16)", "bf16x6": "6 (K=16)", "bf16x1": "1 (K=16)"}
    with torch.no_grad():
        for fmt in ("fp32", "tf32x3", "bf16x6", "bf16x3", "tf32x1", "bf16x1"):
            got = ppo_grads(p, obs, actions, old_logp, values_old, adv, ret, fmt).numpy().astype(np.float64)
            err = np.abs(got - ref)
            print(f"| {fmt} | {cost[fmt]} | {err.max() / scale:.2e} | {np.linalg.norm(got - ref) / np.linalg.norm(ref):.2e} |", flush=True)


if __name__ == "__main__":
    main()
