"""CPU checks of the algebra behind the FP64-mma kernels (pure NumPy, no GPU, no oracle): the moment forms of
csrc/fused_linear.cuh (every estimator term of a linear model from the Gram matrix of the stacked, shifted rows) and the
difference-product identity of csrc/fused_pairs_mma.cuh, against direct per-sample evaluation of the reference's formulas
(src/sobol_sensitivity.jl:192, :197, :207-224)."""
import numpy as np


def test_linear_moment_forms_reproduce_every_estimator_term():
    rng = np.random.default_rng(3)
    d, nout, n = 7, 4, 500
    A, B = rng.random((n, d)) * 3 + 10.0, rng.random((n, d)) * 3 + 10.0        # mean >> spread: the shift matters
    W = rng.normal(size=(nout, d))
    c = A[0]
    a, b = A - c, B - c
    u = np.hstack([a, b])
    U = u.T @ u                                                                 # the Gram matrix the kernel accumulates
    Saa, Sab, Sbb = U[:d, :d], U[:d, d:], U[d:, d:]
    Sba = Sab.T
    sa, sb = a.sum(0), b.sum(0)
    sd = sb - sa
    DD = Sbb - Sba - Sab + Saa
    for o in range(nout):
        w = W[o]
        yA, yB = A @ w, B @ w
        y0 = w @ c
        hB = w @ (Sbb - Sba)                                                    # sum_l w_l sum b'_l delta_k
        hA = w @ (Sab - Saa)
        yb_d, ya_d = y0 * sd + hB, y0 * sd + hA                                 # sum yB delta_k, sum yA delta_k
        sum_ya, sum_ya2 = n * y0 + w @ sa, w @ Saa @ w + 2 * y0 * (w @ sa) + n * y0 * y0
        assert np.isclose(sum_ya, yA.sum(), rtol=1e-13) and np.isclose(sum_ya2, (yA ** 2).sum(), rtol=1e-13)
        s_all = (w @ (sa + sb)) + 2 * n * y0
        q_all = w @ (Saa + Sbb) @ w + 2 * y0 * (w @ (sa + sb)) + 2 * n * y0 * y0
        assert np.isclose(s_all, yA.sum() + yB.sum(), rtol=1e-13) and np.isclose(q_all, (yA ** 2).sum() + (yB ** 2).sum(), rtol=1e-13)
        yAB = yA[:, None] + w[None, :] * (B - A)                                # AB_k: coordinate k from B  (rank-1 update)
        yBA = yB[:, None] - w[None, :] * (B - A)
        scale = np.abs(yA).max() ** 2 * n
        assert np.allclose(w * yb_d, (yB[:, None] * (yAB - yA[:, None])).sum(0), atol=1e-13 * scale)                 # :192
        assert np.allclose(w * w * np.diag(DD), ((yA[:, None] - yAB) ** 2).sum(0), atol=1e-13 * scale)               # :213
        assert np.allclose(-(w * ya_d), (yA[:, None] * (yA[:, None] - yAB)).sum(0), atol=1e-13 * scale)              # :211
        assert np.allclose(sum_ya2 + w * ya_d, (yA[:, None] * yAB).sum(0), rtol=1e-12)                               # :207 / :224
        assert np.allclose(2 * sum_ya2 + 2 * w * ya_d + w * w * np.diag(DD), (yA[:, None] ** 2 + yAB ** 2).sum(0), rtol=1e-12)   # :219
        assert np.allclose(2 * sum_ya + w * sd, (yA[:, None] + yAB).sum(0), rtol=1e-12)                              # :220
        for k in range(d):
            for j in range(k + 1, d):
                want = (yBA[:, k] * yAB[:, j] - yA * yB).sum()                                                       # :197
                got = w[j] * yb_d[j] - w[k] * ya_d[k] - w[k] * w[j] * DD[k, j]
                assert abs(got - want) <= 1e-13 * scale, (o, k, j)


def test_difference_pair_sums_identity():
    """sum (yBA_k yAB_j - yA yB) = V_j + U_k + P_kj with u_k = yBA_k - yB, w_j = yAB_j - yA, V_j = sum yB w_j (the first-order
    numerator, :192), U_k = sum yA u_k and P = sum u w^T (the decomposition the pair kernels accumulate; P is the mma product)."""
    rng = np.random.default_rng(4)
    n, d = 300, 5
    yA, yB = 1.0 + 1e-2 * rng.normal(size=n), 1.0 + 1e-2 * rng.normal(size=n)          # mean >> spread
    yAB, yBA = 1.0 + 1e-2 * rng.normal(size=(n, d)), 1.0 + 1e-2 * rng.normal(size=(n, d))
    u, w = yBA - yB[:, None], yAB - yA[:, None]
    P = u.T @ w                                                                         # what the DMMAs accumulate
    V, U = (yB[:, None] * w).sum(0), (yA[:, None] * u).sum(0)
    for k in range(d):
        for j in range(k + 1, d):
            want = (yBA[:, k] * yAB[:, j] - yA * yB).sum()
            got = P[k, j] + (V[j] + U[k])
            assert abs(got - want) <= 1e-15 * n, (k, j, got, want)
