"""kb_create_multi: several GPUs behind one C-ABI handle.  On a 1-GPU box the same code paths run with
device 0 listed twice (two contexts, two host threads, the peer-to-peer all-reduce kernel on local pointers); on a
multi-GPU box (gpurun --gpus N) the devices are distinct and the all-reduce goes over NVLink."""
import numpy as np
import pytest

from conftest import toeplitz
from oracle import knockoffs_oracle as ko

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 8))) if n > 1 else [0, 0]


@pytest.fixture(scope="module")
def mc(kb):
    m = kb.MultiContext(_devices())
    yield m
    m.close()


def test_multi_gram_equals_single_device(kb, ctx, mc):
    rng = np.random.default_rng(3)
    n, p = 5001, 333                       # odd row count: ragged shards
    X = np.asfortranarray(rng.standard_normal((n, p)) + rng.standard_normal(p))
    G1, mu1 = kb.gram(X, center=True, ctx=ctx)
    G2, mu2 = mc.gram(X, center=True)
    Xc = X - X.mean(0)
    want = Xc.T @ Xc / n
    assert np.max(np.abs(G2 - want)) < 1e-11 and np.max(np.abs(G1 - G2)) < 1e-11
    assert np.max(np.abs(mu2 - X.mean(0))) < 1e-12
    assert np.array_equal(G2, G2.T)
    # dosage data: integer sums are exact whatever the partition
    Xd = kb.harness.dosage_X(4000, 200, 7)
    Gd, _ = mc.gram(Xd, center=False, denom=1.0)
    assert np.array_equal(Gd, Xd.T @ Xd)


@pytest.mark.parametrize("m", [1, 3])
def test_multi_condition_equals_single_device_bit_for_bit(kb, ctx, mc, m):
    rng = np.random.default_rng(4)
    n, p = 3001, 192
    Sigma = toeplitz(p, 0.5)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    mu = rng.standard_normal(p) * 0.1
    s = kb.solve_s(Sigma, "maxent", m=m, ctx=ctx)
    # device RNG keyed by the GLOBAL row: the union of the shards is the single-device result
    a = kb.condition(X, mu, Sigma, s, m=m, seed=11, ctx=ctx)
    b = mc.condition(X, mu, Sigma, s, m=m, seed=11)
    assert np.array_equal(a, b)
    # injected Z against the oracle
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    got = mc.condition(X, mu, Sigma, s, m=m, Z=Z)
    want = ko.condition(X, mu, Sigma, s, m=m, Z=Z)
    assert np.max(np.abs(got - want)) < 1e-10


def test_multi_one_shot_and_solve_many(kb, ctx, mc):
    rng = np.random.default_rng(5)
    n, p, m = 1500, 200, 2
    Sigma = toeplitz(p, 0.4)
    X = np.asfortranarray(rng.standard_normal((n, p)) @ np.linalg.cholesky(Sigma).T)
    Z = np.asfortranarray(rng.standard_normal((n, m * p)))
    res = mc.modelX_gaussian_knockoffs(X, "mvr", np.zeros(p), Sigma, m=m, Z=Z)
    want_X, want_s = ko.modelX_gaussian_knockoffs(X, "mvr", np.zeros(p), Sigma, m=m, Z=Z)
    assert np.max(np.abs(res.s - want_s) / want_s) < 1e-8 and np.max(np.abs(res.Xko - want_X)) < 1e-10
    ss, infos = mc.solve_s_many(Sigma, ["mvr", "sdp_ccd", "maxent"], [m, m, 1])
    for s, (meth, mm) in zip(ss, [("mvr", m), ("sdp_ccd", m), ("maxent", 1)]):
        assert np.max(np.abs(s - kb.solve_s(Sigma, meth, m=mm, ctx=ctx))) == 0.0
    assert infos[1]["sweeps"] == 59


def test_multi_group_blockdiag_equals_full_solve(kb, mc):
    """BASELINE configs[3] logic through the C ABI: LD blocks on different contexts / devices, lockstep sweeps with the
    in-process reduce; the union of the blocks equals the single-context solve of the whole block-diagonal matrix."""
    def block(ng, g, rho, gam):
        grp = np.repeat(np.arange(1, ng + 1), g)
        S = np.where(grp[:, None] == grp[None, :], rho, rho * gam).astype(float)
        np.fill_diagonal(S, 1.0)
        return np.asfortranarray(S), grp
    blocks = [block(6, 5, 0.4, 0.2), block(5, 4, 0.6, 0.3), block(4, 5, 0.5, 0.25)]
    n = sum(b[0].shape[0] for b in blocks)
    full = np.zeros((n, n), order="F")
    gfull = np.zeros(n, dtype=np.int64)
    o = gmax = 0
    for S, g in blocks:
        k = S.shape[0]
        full[o:o + k, o:o + k] = S
        gfull[o:o + k] = g + gmax
        o += k
        gmax += g.max()
    m = 2
    want_S, _, want_obj = kb.solve_s_group(full, gfull, "maxent", m=m, tol=1e-3, ctx=kb.Context(0))
    Sb, obj, infos = mc.solve_s_group_blockdiag([b[0] for b in blocks], [b[1] for b in blocks], "maxent", m=m, tol=1e-3)
    o = 0
    for S in Sb:
        k = S.shape[0]
        assert np.max(np.abs(S - want_S[o:o + k, o:o + k])) < 1e-8
        o += k
    assert abs(obj - want_obj) < 1e-8 * max(1.0, abs(want_obj))
    assert len({i["inner_sweeps"] for i in infos}) == 1


def test_multi_errors(kb, mc):
    with pytest.raises(kb.KnockoffsError):
        kb.MultiContext([])
    with pytest.raises(kb.KnockoffsError):
        kb.MultiContext([99])
    with pytest.raises(kb.KnockoffsError):
        mc.condition(np.zeros((4, 5)), np.zeros(5), np.eye(5), np.full(5, 0.5), m=0)
