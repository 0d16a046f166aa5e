"""Lane-level CPU emulation of the index arithmetic of `ns_resid_f64_kernel` (csrc/agp_umma.cu): shared-memory staging,
mma.sync.m8n8k4.f64 fragment ownership (A[row = lane/4][k = lane%4], B[k = lane%4][n = lane/4], C[row = lane/4][2*(lane%4) + {0,1}],
the mapping agp_tail2.cuh already runs on the GPU) and the epilogue, restated thread by thread; checks T = I - Y P and the
residual sum.  A restatement, so it only proves the arithmetic as written there - it was used once, when the kernel was
written without a GPU at hand.   python tests/studies/ns_resid_kernel_index_emulation.py"""
import numpy as np

m, NSTN, NSK, LD = 128, 32, 32, 36
rng = np.random.default_rng(0)
Y = rng.standard_normal((m, m)).astype(np.float32)
P = rng.standard_normal((m, m))
P = (P + P.T) / 2
T = np.zeros((m, m), np.float32)
res = 0.0
for by in range(m // 64):
    for bx in range(m // NSTN):
        i0, j0 = by * 64, bx * NSTN
        acc = np.zeros((8, 32, NSTN // 8, 2))          # warp, lane, column block, 2
        for k0 in range(0, m, NSK):
            sA, sB = np.zeros((64, LD)), np.zeros((NSTN, LD))
            for t in range(256):
                for u in range(4):
                    e = t + u * 256
                    row, c2 = e >> 4, (e & 15) * 2
                    sA[row, c2:c2 + 2] = Y[i0 + row, k0 + c2:k0 + c2 + 2]
                for u in range(NSTN // 16):
                    e = t + u * 256
                    row, c2 = e >> 4, (e & 15) * 2
                    sB[row, c2:c2 + 2] = P[j0 + row, k0 + c2:k0 + c2 + 2]
            for w in range(8):
                for k in range(0, NSK, 4):
                    Af = np.zeros((8, 4))
                    for lane in range(32):
                        Af[lane >> 2, lane & 3] = sA[8 * w + (lane >> 2), k + (lane & 3)]
                    for nb in range(NSTN // 8):
                        Bf = np.zeros((4, 8))
                        for lane in range(32):
                            Bf[lane & 3, lane >> 2] = sB[8 * nb + (lane >> 2), k + (lane & 3)]
                        Dm = Af @ Bf
                        for lane in range(32):
                            acc[w, lane, nb, 0] += Dm[lane >> 2, 2 * (lane & 3)]
                            acc[w, lane, nb, 1] += Dm[lane >> 2, 2 * (lane & 3) + 1]
        for w in range(8):
            for lane in range(32):
                row = i0 + 8 * w + (lane >> 2)
                for nb in range(NSTN // 8):
                    col = j0 + 8 * nb + 2 * (lane & 3)
                    v0 = (1.0 if row == col else 0.0) - acc[w, lane, nb, 0]
                    v1 = (1.0 if row == col + 1 else 0.0) - acc[w, lane, nb, 1]
                    T[row, col], T[row, col + 1] = v0, v1
                    res += v0 * v0 + v1 * v1
ref = np.eye(m) - Y.astype(np.float64) @ P
err = np.abs(T - ref).max() / np.abs(ref).max()
print(f"max rel deviation of T from I - Y P: {err:.2e};  residual sum {res:.6f} vs {np.sum(ref ** 2):.6f}")
assert err < 1e-6 and abs(res - np.sum(ref ** 2)) < 1e-6 * res
