"""Which parts of a Newton-Schulz refinement  Y <- Y + Y (I - P Y)  need more than fp32?  (CPU emulation, numpy.)
P = identity + decaying spectrum up to cond(P) (the shape of the whitened precision P_v = I + rho V^T theta V), start
20 % off in every direction.  Columns: relative Frobenius error of Y after 1..5 iterations for
  residual in fp32 | fp64,   Y symmetrised after every iteration: no | yes.
    python tests/studies/newton_schulz_precision_study.py
"""
import numpy as np


def spd(m, cond, rng):
    Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    ev = 1 + (cond - 1) * np.exp(-np.linspace(0, 0.1 * m, m))
    return (Q * ev) @ Q.T


for m, cond in [(512, 1e2), (512, 1e4), (1024, 5e4)]:
    for f64res in (0, 1):
        for sym in (0, 1):
            rng = np.random.default_rng(0)
            P = spd(m, cond, rng)
            P = (P + P.T) / 2
            S = np.linalg.inv(P)
            w, Q = np.linalg.eigh(S)
            Sh = (Q * np.sqrt(w)) @ Q.T
            N = rng.standard_normal((m, m))
            N = (N + N.T) / 2
            N /= np.linalg.norm(N, 2)
            Y32 = (Sh @ (np.eye(m) + 0.2 * N) @ Sh).astype(np.float32)
            P32, I = P.astype(np.float32), np.eye(m, dtype=np.float32)
            S32 = np.linalg.inv(P32.astype(np.float64))
            errs = []
            for k in range(5):
                T = (np.eye(m) - Y32.astype(np.float64) @ P32.astype(np.float64)).astype(np.float32) if f64res else I - Y32 @ P32
                Y32 = Y32 + Y32 @ T.T                       # correction product always fp32 (3xTF32 on the tensor cores)
                if sym:
                    Y32 = (Y32 + Y32.T) * np.float32(0.5)
                errs.append(np.linalg.norm(Y32 - S32) / np.linalg.norm(S32))
            print(f"m {m:4d} cond {cond:7.0e}  residual {'fp64' if f64res else 'fp32'}  symmetrise {'yes' if sym else 'no '}  " +
                  "  ".join(f"{e:.1e}" for e in errs))
