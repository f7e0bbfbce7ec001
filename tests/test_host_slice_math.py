"""CPU: the algebra of the shared-memory slice kernel (csrc/slice_eval.cuh), restated in NumPy and checked against the
oracle.  The kernel deals the columns of yTilde to CTAs; every CTA exponentiates against its OWN maximum m_c and
publishes {m_c, S_c, partial prior sums, A_c,i}; after ONE grid barrier the contributions are combined with
s_c = exp(m_c - m).  For the forces method the KL term is formed from K_c = sum_j (x_j - m_c) u_j, i.e. without the
reference's per-structure guard (c_bioen_kernels_forces.c:246-274), which changes it by < 1e-300.  This file pins that
reformulation (not the CUDA code: tests/test_gpu_slice.py does that) -- including column ranges whose maxima lie
hundreds apart, zero prior weights and a single CTA."""
import numpy as np
import pytest

from conftest import grad_err, rel

TOL = 1e-12


def slices(N, nc):
    return [slice(c, min(N, c + nc)) for c in range(0, N, nc)]


def logw_by_slices(g, G, Y, Yobs, theta, nc):
    recs = []
    for sl in slices(g.size, nc):
        x = g[sl]
        m = x.max()
        e = np.exp(x - m)
        recs.append((m, e.sum(), ((x - G[sl]) * e).sum(), (x * e).sum(), (G[sl] * e).sum(), Y[:, sl] @ e, e))
    mx = max(r[0] for r in recs)
    s = [np.exp(r[0] - mx) for r in recs]
    S = sum(r[1] * sc for r, sc in zip(recs, s))
    inv = 1.0 / S
    p0, gbar, Gbar = (sum(r[k] * sc for r, sc in zip(recs, s)) * inv for k in (2, 3, 4))
    avg = sum(r[5] * sc for r, sc in zip(recs, s)) * inv
    r_ = avg - Yobs
    f = theta * (p0 - (mx + np.log(S)) + np.log(np.exp(G).sum())) + 0.5 * float(r_ @ r_)
    w = np.concatenate([r[6] * (sc * inv) for r, sc in zip(recs, s)])
    c = r_ @ (Y - avg[:, None])
    return f, w * theta * (g - gbar - G + Gbar) + w * c, w


def forces_by_slices(fv, w0, Y, Yobs, theta, nc):
    recs = []
    for sl in slices(w0.size, nc):
        x = fv @ Y[:, sl]
        m = x.max()
        u = w0[sl] * np.exp(x - m)
        recs.append((m, u.sum(), ((x - m) * u).sum(), Y[:, sl] @ u, u, x))
    mx = max(r[0] for r in recs)
    s = [np.exp(r[0] - mx) for r in recs]
    S = sum(r[1] * sc for r, sc in zip(recs, s))
    inv, logS = 1.0 / S, np.log(S)
    kl = sum(sc * (r[2] + (r[0] - mx) * r[1]) for r, sc in zip(recs, s)) * inv - logS
    avg = sum(r[3] * sc for r, sc in zip(recs, s)) * inv
    r_ = avg - Yobs
    f = theta * kl + 0.5 * float(r_ @ r_)
    w = np.concatenate([r[4] * (sc * inv) for r, sc in zip(recs, s)])
    x = np.concatenate([r[5] for r in recs])
    tiny = np.finfo(np.float64).tiny
    lr = np.where((w >= tiny) & (w0 >= tiny), x - mx - logS, 0.0)
    E = ((1.0 + lr) * theta + r_ @ Y) * w
    return f, (Y - avg[:, None]) @ E, w


@pytest.mark.parametrize("M,N,nc", [(28, 5001, 34), (100, 777, 22), (5, 20, 8), (7, 3, 8), (808, 10, 8), (17, 9, 8),
                                    (64, 300, 14)])
def test_slice_algebra_matches_the_oracle(oracle, M, N, nc):
    P = oracle.synthetic_problem(M, N, seed=M + 3 * N)
    rng = np.random.default_rng(N)
    Y, Yobs, theta = P["yTilde"], np.asarray(P["YTilde"]).ravel(), 3.7
    G = 0.2 * rng.standard_normal(N)
    g = G + 0.1 * rng.standard_normal(N)
    w0 = rng.random(N) + 0.1
    w0 /= w0.sum()
    fv = 1e-3 * rng.standard_normal(M)
    f, grad, w = logw_by_slices(g, G, Y, Yobs, theta, nc)
    fo, go = oracle.logw_fg(g, G, Y, P["YTilde"], theta)
    assert rel(f, fo) < TOL and grad_err(grad, go) < 1e-11 and abs(w.sum() - 1.0) < 1e-13
    f, grad, w = forces_by_slices(fv, w0, Y, Yobs, theta, nc)
    fo, go = oracle.forces_fg(fv, w0, Y, P["YTilde"], theta)
    assert rel(f, fo) < TOL and grad_err(grad, go) < 1e-11
    assert np.max(np.abs(w - oracle.forces_weights(fv, w0, Y))) < 1e-15


def test_slice_algebra_extreme_maxima_and_zero_prior_weights(oracle):
    M, N, nc, theta = 12, 3000, 42, 2.0
    P = oracle.synthetic_problem(M, N, seed=5)
    rng = np.random.default_rng(6)
    Y, Yobs = P["yTilde"], np.asarray(P["YTilde"]).ravel()
    G = rng.standard_normal(N)
    g = G + rng.standard_normal(N)
    g[:700] += 600.0          # column ranges (CTAs) whose maxima lie 1 200 apart: the far ones underflow to 0, no NaN
    g[700:1500] -= 600.0
    f, grad, w = logw_by_slices(g, G, Y, Yobs, theta, nc)
    fo, go = oracle.logw_fg(g, G, Y, P["YTilde"], theta)
    assert np.isfinite(f) and np.all(np.isfinite(grad))
    assert rel(f, fo) < TOL and grad_err(grad, go) < 1e-11
    w0 = rng.random(N) + 0.1
    w0[100:400] = 0.0         # whole slices without prior weight: S_c = K_c = 0, guarded log-ratio 0
    w0[::7] = 0.0
    w0 /= w0.sum()
    fv = 0.3 * rng.standard_normal(M)
    f, grad, w = forces_by_slices(fv, w0, Y, Yobs, theta, nc)
    fo, go = oracle.forces_fg(fv, w0, Y, P["YTilde"], theta)
    assert np.isfinite(f) and np.all(np.isfinite(grad)) and np.all(w[100:400] == 0.0)
    assert rel(f, fo) < TOL and grad_err(grad, go) < 1e-11
