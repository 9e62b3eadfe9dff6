"""CPU (numpy float32): the two conservativeness properties the ordered / packet traversals rest on,
checked on random boxes and rays with the same float32 arithmetic the kernels use (-fmad=false: every product rounds once).

1. A box the reference's divide-based slab test (accelerators.h:588-626) accepts is accepted by the interior test
   t = plane * rcp(d) with the exit distance widened by 2^-20, for ANY reciprocal within 2^-22 relative of 1/d
   (rcp.approx is within 1 ulp).
2. The packet's hull test — near planes times [min, max] reciprocal, far planes likewise — accepts every box that ANY of the four
   rays' own tests accepts (float multiplication is monotone).
Interior tests only steer (the candidate criterion is leaf-local), so a superset is all exactness needs."""
import numpy as np

F = np.float32
WIDE2 = F(9.53674316e-7)          # 2^-20, traverse.cuh


def ref_slab(d, bmin, bmax):
    """boundingBoxIntersection for origin 0 (accelerators.h:588-626): IEEE divides, x -> y -> z, no t-range test."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (bmin / d).astype(F)
        t1 = (bmax / d).astype(F)
    lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
    tmin, tmax = lo[:, 0].copy(), hi[:, 0].copy()
    ok = ~((tmin > hi[:, 1]) | (lo[:, 1] > tmax))
    tmin = np.where(lo[:, 1] > tmin, lo[:, 1], tmin)
    tmax = np.where(hi[:, 1] < tmax, hi[:, 1], tmax)
    ok &= ~((tmin > hi[:, 2]) | (lo[:, 2] > tmax))
    return ok


def conservative(inv, bmin, bmax):
    """interior test of traverse_fast_loop / traverse_packet (origin 0): products, octant-free min/max, widened exit."""
    a, b = (bmin * inv).astype(F), (bmax * inv).astype(F)
    tmin = np.minimum(a, b).max(axis=1)
    tmax = np.maximum(a, b).min(axis=1)
    tmax = (np.abs(tmax) * WIDE2 + tmax).astype(F)
    return tmin, tmax


def random_boxes(rng, n):
    c = (rng.normal(size=(n, 3)) * [6, 6, 6] + [0, 0, -60]).astype(F)
    h = np.abs(rng.normal(size=(n, 3)) * rng.choice([0.05, 0.5, 5.0], size=(n, 1))).astype(F)
    return (c - h).astype(F), (c + h).astype(F)


def test_reciprocal_interior_test_accepts_what_the_reference_accepts():
    rng = np.random.default_rng(1)
    n = 400000
    bmin, bmax = random_boxes(rng, n)
    # rays aimed near the boxes (grazing cases matter), unit length, generic components
    tgt = (bmin + (bmax - bmin) * rng.uniform(-0.3, 1.3, size=(n, 3))).astype(np.float64)
    d = (tgt / np.linalg.norm(tgt, axis=1, keepdims=True)).astype(F)
    ref = ref_slab(d, bmin, bmax)
    for k in range(3):
        rel = rng.uniform(-1, 1, size=(n, 3)) * 2.0 ** -22
        inv = ((1.0 / d.astype(np.float64)) * (1 + rel)).astype(F)        # any reciprocal within 2^-22 (rcp.approx: 1 ulp)
        tmin, tmax = conservative(inv, bmin, bmax)
        acc = tmin <= tmax
        assert not np.any(ref & ~acc), "the reciprocal-multiply test rejected a box the reference's divide test accepts"
    assert 0.2 < ref.mean() < 0.95            # the sample exercises both outcomes


def test_packet_hull_test_accepts_what_any_ray_accepts():
    rng = np.random.default_rng(2)
    n = 300000
    bmin, bmax = random_boxes(rng, n)
    # four jittered directions per packet, one octant (dz < 0), including packets close to an axis plane
    tgt = (bmin + (bmax - bmin) * rng.uniform(-0.5, 1.5, size=(n, 3))).astype(np.float64)      # aimed near the boxes
    base = (tgt / np.linalg.norm(tgt, axis=1, keepdims=True))[:, None, :]
    base[: n // 10, 0, 0] = rng.uniform(1e-5, 1e-3, n // 10)               # tiny |dx|: wide reciprocal intervals
    jit = rng.normal(size=(n, 4, 3)) * [3e-4, 3e-4, 0]
    d = base + jit
    sx, sy = np.sign(d[:, :1, 0]), np.sign(d[:, :1, 1])
    d[:, :, 0] = np.abs(d[:, :, 0]) * sx
    d[:, :, 1] = np.abs(d[:, :, 1]) * sy
    d = (d / np.linalg.norm(d, axis=2, keepdims=True)).astype(F)
    inv = (F(1) / d).astype(F)
    any_ray = np.zeros(n, bool)
    tlim = np.abs(rng.normal(size=(n, 4)) * 80).astype(F)                  # each ray's own pruning bound
    for j in range(4):
        tmin, tmax = conservative(inv[:, j], bmin, bmax)
        any_ray |= tmin <= np.minimum(tmax, tlim[:, j])
    ilo, ihi = inv.min(axis=1), inv.max(axis=1)
    # hull: near plane -> min over the two products, far plane -> max; octant from the packet's common signs
    neg = inv[:, 0] < 0
    near = np.where(neg, bmax, bmin)
    far = np.where(neg, bmin, bmax)
    tmin_lo = np.minimum((near * ilo).astype(F), (near * ihi).astype(F)).max(axis=1)
    tmax_hi = np.maximum((far * ilo).astype(F), (far * ihi).astype(F)).min(axis=1)
    tmax_hi = (np.abs(tmax_hi) * WIDE2 + tmax_hi).astype(F)
    hull = tmin_lo <= np.minimum(tmax_hi, tlim.max(axis=1))
    assert not np.any(any_ray & ~hull), "the hull test rejected a box one of the packet's rays accepts"
    extra = np.count_nonzero(hull & ~any_ray) / max(1, np.count_nonzero(hull))
    print("hull accepts %.2f %% boxes no single ray accepts (random boxes, not a BVH)" % (100 * extra))
    assert 0.05 < any_ray.mean() < 0.95
