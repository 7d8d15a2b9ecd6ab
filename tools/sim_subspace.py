"""CPU (numpy, fp64) model of the Gram-based subspace iteration: how many power steps / which shift /
which Rayleigh-Ritz window are needed for the 1 % reconstruction-error criterion.  Design aid only."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from xkv_b200 import synthetic  # noqa: E402


def run_case(tokens, cols, rank, alpha, configs):
    x = synthetic.group_matrix(tokens, cols, alpha, seed=1234).double().numpy()
    G = x.T @ x
    t0 = time.time()
    lam = np.linalg.eigvalsh(G)[::-1]
    tr = lam.sum()
    opt = lam[rank:].sum()
    print(f"== tokens={tokens} cols={cols} r={rank} alpha={alpha}: opt err {np.sqrt(opt / tr):.5f}  (eigh {time.time() - t0:.1f}s)"
          f"  lam1/lam_r={lam[0] / lam[rank - 1]:.3g}")
    n = cols
    rng = np.random.default_rng(0)
    for cfg in configs:
        l = -(-(rank + cfg["over"]) // 64) * 64
        l = min(l, n)
        q = cfg["q"]
        W = cfg.get("W", 128)
        shift_mode = cfg.get("shift", None)
        om = rng.standard_normal((n, l))
        Y = G @ om
        Q, R = np.linalg.qr(Y)
        for it in range(q):
            c = 0.0
            if shift_mode is not None:
                # estimate of lam_l from the Rayleigh quotients of the trailing basis vectors
                d = np.einsum("ij,ij->j", Q[:, -8:], G @ Q[:, -8:])
                c = shift_mode * d.mean()
            Y = G @ Q - c * Q
            Q, R = np.linalg.qr(Y)
        # windowed Rayleigh-Ritz around column r
        out = {}
        for Wt in ([W] if not isinstance(W, (list, tuple)) else W):
            V = Q.copy()
            if Wt > 0:
                wr = l - rank
                wl = min(rank, Wt - wr)
                if wl > 0:
                    r0 = rank - wl
                    Qw = Q[:, r0:l]
                    T = Qw.T @ G @ Qw
                    ev, Z = np.linalg.eigh(T)
                    Z = Z[:, ::-1]
                    V[:, r0:rank] = Qw @ Z[:, :wl]
            Vr = V[:, :rank]
            cap = np.trace(Vr.T @ G @ Vr)
            out[Wt] = np.sqrt((tr - cap) / opt)
        # full RR for comparison
        T = Q.T @ G @ Q
        ev = np.linalg.eigvalsh(T)[::-1]
        full = np.sqrt((tr - ev[:rank].sum()) / opt)
        print(f"   over={cfg['over']:4d} l={l:4d} q={q} shift={shift_mode}: ratio windowed " +
              " ".join(f"W{k}:{v:.5f}" for k, v in out.items()) + f"   full-RR {full:.5f}")


if __name__ == "__main__":
    cases = [
        (4096, 4096, 512, 1.0),
        (4096, 4096, 768, 0.5),
        (4096, 4096, 512, None),
        (1536, 2048, 512, 1.0),
    ]
    configs = []
    for q in (1, 2, 3, 4, 6):
        configs.append(dict(over=64, q=q, W=[0, 128, 256]))
    for q in (1, 2, 3, 4):
        configs.append(dict(over=64, q=q, W=[0, 128, 256], shift=0.5))
    for q in (1, 2, 3):
        configs.append(dict(over=192, q=q, W=[0, 256, 384]))
    only = int(sys.argv[1]) if len(sys.argv) > 1 else None
    for i, c in enumerate(cases):
        if only is None or only == i:
            run_case(*c, configs)
