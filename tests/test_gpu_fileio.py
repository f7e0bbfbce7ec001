"""GPU: matrices streamed from disk into HBM in row chunks (bioen_b200.fileio.upload_streamed) and reference-format
problem files evaluated straight from the file."""
import numpy as np
import pytest

from conftest import grad_err, rel

pytestmark = pytest.mark.gpu


def test_streamed_upload_equals_direct_upload(oracle, tmp_path):
    import bioen_b200
    from bioen_b200 import fileio as fio
    M, N = 301, 5003
    P = oracle.synthetic_problem(M, N, seed=3)
    fn = str(tmp_path / "yTilde.npy")
    np.save(fn, P["yTilde"])
    g1 = 0.1 * np.random.default_rng(1).standard_normal(N)
    with bioen_b200.Problem(P["yTilde"]) as direct:
        direct.set_logw(P["G"], P["YTilde"], 2.0)
        f0, g0 = direct.objective_and_gradient(g1)
    for chunk_bytes in (8 * N * 7, 8 * N * 64, 1 << 30):          # 43 ragged chunks, 5 chunks, one chunk
        p = fio.upload_streamed(fn, chunk_bytes=chunk_bytes)
        try:
            assert np.array_equal(p.download(), P["yTilde"])
            p.set_logw(P["G"], P["YTilde"], 2.0)
            f, g = p.objective_and_gradient(g1)
            assert f == f0 and np.array_equal(g, g0)
        finally:
            p.close()
    # float32 on disk is converted on the way in; a Fortran-ordered source as well
    np.save(fn, np.asfortranarray(P["yTilde"].astype(np.float32)))
    with fio.upload_streamed(fn, chunk_bytes=8 * N * 50) as p:
        assert np.array_equal(p.download(), P["yTilde"].astype(np.float32).astype(np.float64))
    with pytest.raises(ValueError):
        with bioen_b200.Problem(shape=(M, N + 1)) as q:
            fio.upload_streamed(fn, problem=q)


def test_problem_file_to_optimum(oracle, tmp_path):
    from bioen_b200 import fileio as fio
    P = oracle.synthetic_problem(40, 3000, seed=9)
    fn = str(tmp_path / "problem.pkl")
    fio.dump(fn, [P["GInit"], P["G"], P["y"], P["yTilde"], P["YTilde"], P["w0"], 5.0])
    prob, d = fio.problem_from_file(fn)
    with prob:
        prob.set_logw(d["G"], d["YTilde"], d["theta"])
        f, g = prob.objective_and_gradient(d["GInit"])
        fo, go = oracle.logw_fg(P["GInit"], P["G"], P["yTilde"], P["YTilde"], 5.0)
        assert rel(f, fo) < 1e-11 and grad_err(g, go) < 1e-11
