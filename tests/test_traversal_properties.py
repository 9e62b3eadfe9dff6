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


def ref_slab_origin(o, d, bmin, bmax):
    """boundingBoxIntersection with a ray origin (accelerators.h:588-626): fl(fl(plane - o) / d), x -> y -> z, no t-range test."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t0 = ((bmin - o).astype(F) / d).astype(F)
        t1 = ((bmax - o).astype(F) / d).astype(F)
    lo, hi = np.minimum(t0, t1), np.maximum(t0, t1)
    tmin, tmax = lo[:, 0].copy(), hi[:, 0].copy()
    ok = ~((tmin > hi[:, 1]) | (lo[:, 1] > tmax))
    tmin = np.where(lo[:, 1] > tmin, lo[:, 1], tmin)
    tmax = np.where(hi[:, 1] < tmax, hi[:, 1], tmax)
    ok &= ~((tmin > hi[:, 2]) | (lo[:, 2] > tmax))
    return ok


def fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64, one rounding of the sum to float64 and one to
    float32 (double rounding can differ from a true FMA by an ulp in rare ties - far inside the widening under test)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def affine_interior(o, inv, bmin, bmax, rb_min, rb_max):
    """interior test of rays with a non-zero origin (traverse.cuh, ray_affine + the !ZERO_O branch of the traversal loops):
    t = plane * (1/d) + c, c = -(o * 1/d) -/+ w, w = (max |plane| over the root box + |o|) * |1/d| * 2^-20."""
    k = F(2.0 ** -20)
    R = np.maximum(np.abs(rb_min), np.abs(rb_max)).astype(F)
    w = (((R + np.abs(o)).astype(F) * np.abs(inv)).astype(F) * k).astype(F)
    c = (-(o * inv).astype(F)).astype(F)
    cn, cf = (c - w).astype(F), (c + w).astype(F)
    neg = inv < 0
    near, far = np.where(neg, bmax, bmin), np.where(neg, bmin, bmax)
    tmin = fma32(near, inv, cn).max(axis=1)
    tmax = fma32(far, inv, cf).min(axis=1)
    return tmin, tmax, w.max(axis=1)


def test_affine_interior_test_accepts_what_the_reference_accepts():
    """Round 2: shadow / secondary rays test interior boxes with ONE FFMA per plane. The absolute widening folded into the per-ray
    constants must cover the cancellation in plane * i - o * i: every box the reference's fl(fl(plane - o) / d) test accepts is
    accepted - origins on other boxes' surfaces (shadow rays), far outside the scene, direction components down to 1e-12, boxes
    that contain the origin (t around 0), any reciprocal within 2^-22 of 1/d."""
    rng = np.random.default_rng(3)
    n = 400000
    bmin, bmax = random_boxes(rng, n)
    rb_min, rb_max = bmin.min(axis=0), bmax.max(axis=0)                      # the root box of the "scene"
    # origins: on / near another box (shadow rays leave a surface), some inside the tested box, some far outside the scene
    other = rng.permutation(n)
    o = (bmin[other] + (bmax[other] - bmin[other]) * rng.uniform(-0.1, 1.1, size=(n, 3))).astype(F)
    inside = rng.random(n) < 0.15
    o[inside] = (bmin[inside] + (bmax[inside] - bmin[inside]) * rng.uniform(0, 1, size=(inside.sum(), 3))).astype(F)
    far_o = rng.random(n) < 0.1
    o[far_o] = (o[far_o] * rng.choice([30.0, 1e3, 1e5], size=(far_o.sum(), 1))).astype(F)
    # directions aimed near the tested box from the origin (grazing cases), some with one tiny component
    tgt = (bmin + (bmax - bmin) * rng.uniform(-0.3, 1.3, size=(n, 3))).astype(np.float64)
    # half of the rays aim EXACTLY at a point on an edge of the box: entry and exit distance then coincide up to rounding and the
    # reference's answer hangs on single ulps (without the widening the FFMA form rejects ~14 % of the boxes the reference accepts
    # on such rays; checked when this test was written)
    edge = rng.random(n) < 0.5
    corner = np.where(rng.integers(0, 2, size=(n, 3)).astype(bool), bmin, bmax).astype(np.float64)
    free = rng.integers(0, 3, size=n)
    idx = np.arange(n)
    corner[idx, free] = bmin[idx, free] + (bmax[idx, free] - bmin[idx, free]) * rng.uniform(0, 1, size=n)
    tgt[edge] = corner[edge]
    tgt = tgt - o
    tgt[np.abs(tgt) < 1e-9] = 1e-9
    d = tgt / np.linalg.norm(tgt, axis=1, keepdims=True)
    tiny = rng.random(n) < 0.15
    ax = rng.integers(0, 3, size=n)
    d[tiny, ax[tiny]] = rng.choice([1e-5, 1e-8, 1e-12], size=tiny.sum()) * rng.choice([-1, 1], size=tiny.sum())
    d = d.astype(F)
    ref = ref_slab_origin(o, d, bmin, bmax)
    for k in range(3):
        rel = rng.uniform(-1, 1, size=(n, 3)) * 2.0 ** -22
        inv = ((1.0 / d.astype(np.float64)) * (1 + rel)).astype(F)
        tmin, tmax, wmax = affine_interior(o, inv, bmin, bmax, rb_min, rb_max)
        fast = (np.abs(inv).min(axis=1) > 1e-30) & (np.abs(inv).max(axis=1) < 1e30) & (wmax < 1e30)      # else: the divide-based traversal
        acc = tmin <= tmax
        bad = ref & fast & ~acc
        assert not np.any(bad), "the FFMA interior test rejected %d boxes the reference's divide test accepts" % bad.sum()
    assert fast.mean() > 0.95 and 0.2 < ref.mean() < 0.98
    # and it is not vacuous: boxes the reference rejects are still rejected most of the time, and WITHOUT the widening the same
    # arithmetic does reject boxes the reference accepts (the sample contains the hard cases)
    assert (~acc & ~ref).sum() > 0.5 * (~ref).sum()
    c0 = (-(o * inv).astype(F)).astype(F)
    neg = inv < 0
    t0 = fma32(np.where(neg, bmax, bmin), inv, c0).max(axis=1)
    t1 = fma32(np.where(neg, bmin, bmax), inv, c0).min(axis=1)
    assert np.count_nonzero(ref & fast & ~(t0 <= t1)) > 1000
