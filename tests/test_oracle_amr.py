"""AMR operators: the C oracle restatement against the reference's OWN
structs -- parthenon::refinement_ops::RestrictAverage (prolong_restrict/pr_ops.hpp:105-165) and
Hydro::refinement_ops::ProlongateCellMinModMultiD (src/hydro/prolongation/custom_ops.hpp:49-186)
compiled unmodified into oracle/_ref -- bit for bit on random data."""
import numpy as np
import pytest

import oracle_lib
import ref_lib

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")


def shapes(ndim, nx, ng):
    nf = [nx + 2 * ng if d < ndim else 1 for d in range(3)]
    nc = [nx // 2 + 2 * ng if d < ndim else 1 for d in range(3)]
    return nf, nc


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("el", [0, 1, 2, 3])
def test_restrict_average_matches_reference(ndim, el):
    if el > ndim:
        pytest.skip("no such face direction")
    rng = np.random.default_rng(10 * ndim + el)
    nx, ng, nv = 8, 3, 9
    nf, nc = shapes(ndim, nx, ng)
    fine = rng.standard_normal((nv, nf[2], nf[1], nf[0]))
    xmin, dx = (0.1, -0.3, 0.7), (0.013, 0.017, 0.011)
    # interior coarse cells (what SendBoundBufs restricts lie inside this range)
    lo = [ng if d < ndim else 0 for d in range(3)]
    hi = [ng + nx // 2 - 1 if d < ndim else 0 for d in range(3)]
    if el > 0:
        hi[el - 1] += 1  # faces: one more along the normal
        hi[el - 1] = min(hi[el - 1], nc[el - 1] - 1)
    a = np.full((nv, nc[2], nc[1], nc[0]), np.nan)
    b = a.copy()
    oracle_lib.restrict_average(ndim, ng, el, xmin, dx, fine, a, lo, hi)
    ref_lib.restrict_average(ndim, ng, el, xmin, dx, fine, b, lo, hi)
    assert np.array_equal(a, b, equal_nan=True)
    assert np.isfinite(a).sum() == nv * np.prod([h - l + 1 for l, h in zip(lo, hi)])


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("kind", ["smooth", "rough", "positive"])
def test_prolongate_minmod_multid_matches_reference(ndim, kind):
    rng = np.random.default_rng(100 * ndim + len(kind))
    nx, ng, nv = 8, 3, 5
    nf, nc = shapes(ndim, nx, ng)
    coarse = rng.standard_normal((nv, nc[2], nc[1], nc[0]))
    if kind == "smooth":
        coarse = np.cumsum(np.abs(coarse), axis=3) * 0.1
    if kind == "positive":
        coarse = np.exp(coarse)
    coarse = np.ascontiguousarray(coarse)
    xmin, dx = (0.25, 1.5, -2.0), (1.0 / 64, 1.0 / 48, 1.0 / 80)
    # every coarse cell that has a full 3-point stencil
    lo = [1 if d < ndim else 0 for d in range(3)]
    hi = [nc[d] - 2 if d < ndim else 0 for d in range(3)]
    # fine index (c - ng) * 2 + ng must stay inside the fine array: restrict to the valid range
    for d in range(ndim):
        lo[d] = max(lo[d], ng - ng // 2)
        hi[d] = min(hi[d], (nf[d] - 2 - ng) // 2 + ng)
    a = np.full((nv, nf[2], nf[1], nf[0]), np.nan)
    b = a.copy()
    oracle_lib.prolongate_minmod_multid(ndim, ng, xmin, dx, a, coarse, lo, hi)
    ref_lib.prolongate_minmod_multid(ndim, ng, xmin, dx, b, coarse, lo, hi)
    assert np.array_equal(a, b, equal_nan=True)
    assert np.isfinite(a).any()
    # no new extrema (the point of the limiter): fine values stay within the coarse neighbourhood
    assert np.nanmax(a) <= coarse.max() + 1e-12 and np.nanmin(a) >= coarse.min() - 1e-12


def test_restriction_inverts_prolongation_of_linear_data():
    """size-independent property: restricting the prolongation returns the coarse data (the
    prolongation is conservative: the 2^d children average to the parent)."""
    ndim, nx, ng, nv = 3, 8, 3, 2
    nf, nc = shapes(ndim, nx, ng)
    rng = np.random.default_rng(7)
    coarse = np.ascontiguousarray(rng.standard_normal((nv, nc[2], nc[1], nc[0])))
    xmin, dx = (0.0, 0.0, 0.0), (0.1, 0.1, 0.1)
    lo, hi = [ng] * 3, [ng + nx // 2 - 1] * 3
    fine = np.full((nv, nf[2], nf[1], nf[0]), np.nan)
    oracle_lib.prolongate_minmod_multid(ndim, ng, xmin, dx, fine, coarse, lo, hi)
    back = np.full_like(coarse, np.nan)
    oracle_lib.restrict_average(ndim, ng, 0, xmin, dx, fine, back, lo, hi)
    sl = (slice(None),) + tuple(slice(ng, ng + nx // 2) for _ in range(3))
    assert np.allclose(back[sl], coarse[sl], rtol=0, atol=1e-13)


def test_pressure_gradient_tag_thresholds():
    ndim, ng = 3, 2
    n = (12, 12, 12)
    prim = np.ones((5, n[2], n[1], n[0]))
    tag, m = oracle_lib.pressure_gradient_tag(ndim, ng, prim, 4, 0.5)
    assert (tag, m) == (-1, 0.0)
    prim[4, 6, 6, 7] = 3.0
    tag, m = oracle_lib.pressure_gradient_tag(ndim, ng, prim, 4, 0.5)
    assert tag == 1 and m == 1.0   # 0.5 * (3 - 1) / 1
    tag, _ = oracle_lib.pressure_gradient_tag(ndim, ng, prim, 4, 3.0)
    assert tag == 0


@pytest.mark.parametrize("el", [1, 2, 3])
def test_flux_restriction_is_conservative(el):
    """Flux correction (bvals_in_one.hpp:69-80): the restricted coarse-face flux times the coarse
    face area equals the sum of the fine-face fluxes times the fine face areas, so replacing the
    coarse flux by it makes the coarse and fine updates see the same mass through the shared face."""
    ndim, nx, ng, nv = 3, 8, 3, 5
    nf, nc = shapes(ndim, nx, ng)
    rng = np.random.default_rng(el)
    fine = rng.standard_normal((nv, nf[2], nf[1], nf[0]))
    dx = (0.02, 0.03, 0.05)
    lo, hi = [ng] * 3, [ng + nx // 2 - 1] * 3
    hi[el - 1] += 1
    coarse = np.full((nv, nc[2], nc[1], nc[0]), np.nan)
    oracle_lib.restrict_average(ndim, ng, el, (0.0, 0.0, 0.0), dx, fine, coarse, lo, hi)
    t = [d for d in range(3) if d != el - 1]          # transverse directions of the face
    fine_area = dx[t[0]] * dx[t[1]]
    for ck in range(lo[2], hi[2] + 1):
        for cj in range(lo[1], hi[1] + 1):
            for ci in range(lo[0], hi[0] + 1):
                c = [ci, cj, ck]
                f0 = [(c[d] - ng) * 2 + ng for d in range(3)]
                tot = 0.0
                for a in (0, 1):
                    for b in (0, 1):
                        f = list(f0)
                        f[t[0]] += a
                        f[t[1]] += b
                        tot = tot + fine[:, f[2], f[1], f[0]] * fine_area
                assert np.allclose(coarse[:, ck, cj, ci] * (4 * fine_area), tot, rtol=1e-14, atol=1e-16)
