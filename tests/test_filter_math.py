"""The error-bound design of the shadow filter (csrc/dev_shadow.cuh), checked off-device: a NumPy float32 mirror of
filter_sphere / filter_plane / filter_cube against a float32 mirror of the REFERENCE's arithmetic (sphere.rs:47-70,
plane.rs:45-56, cube.rs:55-129 behind world.rs:104-119: normalised direction, every root divided out, `t >= 0`,
`t < distance`) on millions of segments concentrated where the two could disagree — tangent rays, origins and light
points on the surface, near-parallel planes, slab corners.  The property: whenever the filter answers HIT or MISS the
reference's boolean is the same; it may answer UNSURE as often as it likes (rarely, away from those places).

This mirrors the MATH (same operations, same order; fused multiply-adds emulated exactly in float64, the approximate
reciprocal / rsqrt perturbed by up to 2 ulp); that the CUDA code computes the same booleans as the exact test is what
the GPU tests check frame by frame (filter on / off equality, 31 scenes)."""
import numpy as np
import pytest

F = np.float32
MISS, HIT, UNSURE = 0, 1, 2
K_ACNE = F(1.1920929e-7) * F(10000.0)
TOL_P = F(3.814697265625e-06)


def fma(a, b, c):  # exact for float32 inputs: the product fits a float64 mantissa
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def approx(x, rng):  # MUFU.RCP / MUFU.RSQ results: correctly rounded value off by up to 2 ulp
    x = x.astype(F)
    return x * (F(1.0) + (rng.integers(-2, 3, size=x.shape).astype(F) * F(2.0 ** -23)))


def xf_vec(m, v):  # matrix.rs:73-84, left to right, no fusing (w = 0)
    return np.stack([(m[:, r, 0] * v[:, 0] + m[:, r, 1] * v[:, 1]) + m[:, r, 2] * v[:, 2] for r in range(3)], axis=1)


def xf_point(m, t, p):
    return np.stack([((m[:, r, 0] * p[:, 0] + m[:, r, 1] * p[:, 1]) + m[:, r, 2] * p[:, 2]) + t[:, r] for r in range(3)], axis=1)


def reference_direction(p, light):
    v = light - p
    dist = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    return v, dist, v / dist[:, None]


def reference_sphere(m, o, p, light):
    """World::is_shadowed against one sphere: the lowest t >= 0 of the two roots, shadowed iff t < distance."""
    v, dist, d = reference_direction(p, light)
    d2 = xf_vec(m, d)
    a = (d2[:, 0] * d2[:, 0] + d2[:, 1] * d2[:, 1]) + d2[:, 2] * d2[:, 2]
    b = F(2.0) * ((d2[:, 0] * o[:, 0] + d2[:, 1] * o[:, 1]) + d2[:, 2] * o[:, 2])
    c = ((o[:, 0] * o[:, 0] + o[:, 1] * o[:, 1]) + o[:, 2] * o[:, 2]) - F(1.0)
    disc = b * b - F(4.0) * a * c
    with np.errstate(invalid="ignore", divide="ignore"):
        ds = np.sqrt(disc)
        t0, t1 = (-b - ds) / (F(2.0) * a), (-b + ds) / (F(2.0) * a)
    t = np.where(t0 >= 0, t0, np.where(t1 >= 0, t1, F(-1.0)))
    # stable sort by t then first t >= 0 (intersection.rs:30-35): with t0 > t1 (a < 0 never happens) the same
    return (disc >= 0) & (t >= 0) & (t < dist)


def filter_sphere(m, o, p, light, tol, rng):
    v = light - p
    d = [fma(m[:, r, 0], v[:, 0], fma(m[:, r, 1], v[:, 1], m[:, r, 2] * v[:, 2])) for r in range(3)]
    a = fma(d[0], d[0], fma(d[1], d[1], d[2] * d[2]))
    b = fma(d[0], o[:, 0], fma(d[1], o[:, 1], d[2] * o[:, 2]))
    oo = fma(o[:, 0], o[:, 0], fma(o[:, 1], o[:, 1], o[:, 2] * o[:, 2]))
    c = oo - F(1.0)
    disc = fma(b, b, -(a * c))
    spread = oo + np.abs(c)
    td = tol * (a * spread)
    out = np.full(len(a), UNSURE)
    out[disc < -td] = MISS
    go = disc > td
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        rs, ia = approx(F(1.0) / np.sqrt(disc), rng), approx(F(1.0) / a, rng)
        sq = disc * rs
        s0, s1 = (-b - sq) * ia, (-b + sq) * ia
        x = oo * ia
        e = tol * (x * approx(F(1.0) / np.sqrt(x + F(1e-30)), rng) + spread * rs + np.maximum(np.abs(s0), np.abs(s1)) + F(1.0))
        p0, n0, p1, n1 = s0 > e, s0 < -e, s1 > e, s1 < -e
        cand = np.where(p0, s0, s1)
        res = np.full(len(a), UNSURE)
        decided = p0 | (n0 & p1)
        res[decided & (cand < F(1.0) - e)] = HIT
        res[decided & (cand > F(1.0) + e)] = MISS
        res[n0 & n1] = MISS
    out[go] = res[go]
    return out


def random_transforms(n, rng, max_cond=64.0):
    """Inverse transforms of spheres: rotation * shear * non-uniform scale, condition number (inf-norm) <= max_cond."""
    def rot(axis, ang):
        c, s = np.cos(ang), np.sin(ang)
        r = np.tile(np.eye(3), (n, 1, 1))
        i, j = [(1, 2), (0, 2), (0, 1)][axis]
        r[:, i, i], r[:, j, j], r[:, i, j], r[:, j, i] = c, c, -s, s
        return r
    scale = np.exp(rng.uniform(np.log(0.15), np.log(2.5), size=(n, 3)))
    squash = rng.random(n) < 0.5
    scale[~squash] = scale[~squash, :1]
    shear = np.tile(np.eye(3), (n, 1, 1))
    shear[:, 0, 1] = np.where(rng.random(n) < 0.3, rng.uniform(-0.6, 0.6, n), 0.0)
    fwd = rot(1, rng.uniform(0, 6.28, n)) @ rot(2, rng.uniform(0, 6.28, n)) @ shear @ (np.eye(3)[None] * scale[:, None, :])
    plain = rng.random(n) < 0.4  # translation * uniform scaling, like the demo scenes
    fwd[plain] = np.eye(3)[None] * scale[plain, :1, None]
    inv = np.linalg.inv(fwd)
    cond = np.abs(inv).sum(axis=2).max(axis=1) * np.abs(fwd).sum(axis=2).max(axis=1)
    keep = cond <= max_cond
    centre = rng.uniform(-4, 4, size=(n, 3))
    t = -(inv @ centre[:, :, None])[:, :, 0]
    return inv[keep].astype(F), t[keep].astype(F), fwd[keep], centre[keep], cond[keep]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sphere_filter_never_contradicts_the_reference(seed):
    rng = np.random.default_rng(seed)
    inv, t, fwd, centre, cond = random_transforms(400_000, rng)
    n = len(inv)
    tol = F(2.0 ** -24 * (128.0 * float(cond.max()) + 32.0))  # SmallScene::tol_sphere for the worst sphere of a scene
    # a tangent segment: a point on the sphere, a tangent direction, then both ends pushed off by tiny amounts
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(u, rng.normal(size=(n, 3)))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    kind = rng.integers(0, 5, n)
    eps = 10.0 ** rng.uniform(-9, -1, n) * rng.choice([-1.0, 1.0], n)
    lift = np.where(kind == 0, 1.0 + eps, 1.0 + np.abs(rng.normal(0, 0.5, n)))             # kind 0: near-tangent
    a_obj = u * lift[:, None] - w * rng.uniform(0.5, 6, (n, 1))
    b_obj = u * lift[:, None] + w * rng.uniform(0.5, 6, (n, 1))
    a_obj = np.where((kind == 1)[:, None], u * (1.0 + eps)[:, None], a_obj)                # kind 1: origin on the surface
    b_obj = np.where((kind == 2)[:, None], u * (1.0 + eps)[:, None], b_obj)                # kind 2: light on the surface
    rand = kind >= 3                                                                        # kinds 3, 4: anything
    a_obj[rand] = rng.normal(0, 2.5, (rand.sum(), 3))
    b_obj[rand] = rng.normal(0, 2.5, (rand.sum(), 3))
    p = ((fwd @ a_obj[:, :, None])[:, :, 0] + centre).astype(F)
    light = ((fwd @ b_obj[:, :, None])[:, :, 0] + centre).astype(F)
    o = xf_point(inv, t, p)  # the reference's object-space origin: shared by both
    want = reference_sphere(inv, o, p, light)
    got = filter_sphere(inv, o, p, light, tol, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), (int(wrong.sum()), np.flatnonzero(wrong)[:5], cond[wrong][:5])
    decided = got != UNSURE
    assert decided[rand].mean() > 0.97, decided[rand].mean()   # ordinary segments are decided
    assert decided[kind == 0].mean() > 0.15                     # near-tangent ones (offsets down to 1e-9) mostly are not


def reference_plane(r1, p, light):
    v, dist, d = reference_direction(p, light)
    oy = ((r1[:, 0] * p[:, 0] + r1[:, 1] * p[:, 1]) + r1[:, 2] * p[:, 2]) + r1[:, 3]
    dy = (r1[:, 0] * d[:, 0] + r1[:, 1] * d[:, 1]) + r1[:, 2] * d[:, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        t = -oy / dy
    return oy, ~(np.abs(dy) < K_ACNE) & (t >= 0) & (t < dist)


def filter_plane(r1, oy, p, light, rng):
    v = light - p
    vv = fma(v[:, 0], v[:, 0], fma(v[:, 1], v[:, 1], v[:, 2] * v[:, 2]))
    length = vv * approx(F(1.0) / np.sqrt(vv), rng)
    px, py, pz = r1[:, 0] * v[:, 0], r1[:, 1] * v[:, 1], r1[:, 2] * v[:, 2]
    dy = px + py + pz
    edy = TOL_P * (np.abs(px) + np.abs(py) + np.abs(pz))
    mag, thr = np.abs(dy), K_ACNE * length
    out = np.full(len(dy), UNSURE)
    steep = (mag - edy > thr * (F(1.0) + TOL_P)) & (oy != 0)
    out[mag + edy < thr * (F(1.0) - TOL_P)] = MISS
    away = steep & ((oy < 0) == (dy < 0))
    out[away] = MISS
    toward = steep & ~away
    aoy = np.abs(oy)
    out[toward & (aoy < (mag - edy) * (F(1.0) - TOL_P))] = HIT
    out[toward & (aoy > (mag + edy) * (F(1.0) + TOL_P))] = MISS
    return out


@pytest.mark.parametrize("seed", [4, 5])
def test_plane_filter_never_contradicts_the_reference(seed):
    rng = np.random.default_rng(seed)
    n = 600_000
    normal = rng.normal(size=(n, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    normal *= np.exp(rng.uniform(np.log(0.1), np.log(10), (n, 1)))  # scaled planes
    r1 = np.concatenate([normal, rng.uniform(-3, 3, (n, 1))], axis=1).astype(F)
    p = rng.normal(0, 3, (n, 3))
    light = rng.normal(0, 3, (n, 3))
    kind = rng.integers(0, 4, n)
    nn = normal / (np.linalg.norm(normal, axis=1, keepdims=True) ** 2)
    height = lambda q: (q * normal).sum(axis=1) + r1[:, 3]  # noqa: E731
    eps = 10.0 ** rng.uniform(-9, -2, n) * rng.choice([-1.0, 1.0], n)
    # kind 0: light (almost) on the plane; kind 1: origin (almost) on it; kind 2: segment (almost) parallel to it
    light = np.where((kind == 0)[:, None], light - nn * (height(light) - eps)[:, None], light)
    p = np.where((kind == 1)[:, None], p - nn * (height(p) - eps)[:, None], p)
    para = kind == 2
    light[para] = (light - nn * (height(light) - height(p) * (1.0 + eps * 50))[:, None])[para]
    p, light = p.astype(F), light.astype(F)
    oy, want = reference_plane(r1, p, light)
    got = filter_plane(r1, oy, p, light, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), (int(wrong.sum()), np.flatnonzero(wrong)[:5])
    assert (got != UNSURE)[kind == 3].mean() > 0.99


def reference_cube(scale, t, p, light):
    """cube.rs:55-63 + aabb_intersection (cube.rs:90-129) for a diagonal inverse: hit at lo if lo >= 0 else at hi."""
    v, dist, d = reference_direction(p, light)
    o = scale * p + t  # one product and one sum per component: exactly xf_point for a diagonal matrix
    d2 = scale * d
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        inv = F(1.0) / d2
        lo_k, hi_k = (F(-1.0) - o) * inv, (F(1.0) - o) * inv
        near, far = np.fmin(lo_k, hi_k), np.fmax(lo_k, hi_k)  # f32::min / max: the non-NaN operand
        lo, hi = near[:, 0], far[:, 0]
        for k in (1, 2):
            lo, hi = np.fmax(lo, near[:, k]), np.fmin(hi, far[:, k])
        hit = hi >= np.fmax(F(0.0), lo)
        tt = np.where(lo >= 0, lo, hi)
    return o, hit & (tt >= 0) & (tt < dist)


def filter_cube(scale, o, p, light, rng):
    v = light - p
    d = scale * v
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        inv = approx(F(1.0) / d, rng)
        a, b = (F(-1.0) - o) * inv, (F(1.0) - o) * inv
        near, far = np.fmin(a, b), np.fmax(a, b)
        lo, hi = near[:, 0], far[:, 0]
        for k in (1, 2):
            lo, hi = np.fmax(lo, near[:, k]), np.fmin(hi, far[:, k])
        e = TOL_P * (np.abs(lo) + np.abs(hi) + F(1.0))
        g = hi - np.fmax(lo, F(0.0))
        out = np.full(len(lo), UNSURE)
        out[g < -e] = MISS
        ok = (g > e) & (np.abs(lo) > e) & (d != 0).all(axis=1)
        cand = np.where(lo > 0, lo, hi)
        out[ok & (cand < F(1.0) - e)] = HIT
        out[ok & (cand > F(1.0) + e)] = MISS
        out[(d == 0).any(axis=1)] = UNSURE
    return out


@pytest.mark.parametrize("seed", [6, 7])
def test_cube_filter_never_contradicts_the_reference(seed):
    rng = np.random.default_rng(seed)
    n = 600_000
    half = np.exp(rng.uniform(np.log(0.01), np.log(3), (n, 3)))  # down to the lampshade's 100:1 slab
    centre = rng.uniform(-3, 3, (n, 3))
    scale, t = (1.0 / half).astype(F), (-centre / half).astype(F)
    p = rng.normal(0, 3, (n, 3))
    light = rng.normal(0, 3, (n, 3))
    kind = rng.integers(0, 4, n)
    eps = 10.0 ** rng.uniform(-9, -2, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
    face = centre + half * np.where(rng.random((n, 3)) < 0.4, rng.choice([-1.0, 1.0], (n, 3)), rng.uniform(-1, 1, (n, 3)))
    light = np.where((kind == 0)[:, None], face + eps * half, light)   # light on a face / edge / corner
    p = np.where((kind == 1)[:, None], face + eps * half, p)           # origin on a face / edge / corner
    graze = kind == 2                                                  # segment through an edge region
    light[graze] = (face + (face - p) * rng.uniform(0.1, 2, (n, 1)) + eps * half)[graze]
    p, light = p.astype(F), light.astype(F)
    o, want = reference_cube(scale, t, p, light)
    got = filter_cube(scale, o, p, light, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), (int(wrong.sum()), np.flatnonzero(wrong)[:5])
    assert (got != UNSURE)[kind == 3].mean() > 0.99


# ---------------------------------------------------------------------------------------------------
# Scale extremes: everything above lives in a +-10-unit world of unit-scale objects.  The bounds are
# RELATIVE (see DESIGN.md "Shadow filter: where the bounds come from"), so they must hold for a millimetre world, a
# kilometre world, a light 10^4 units away, spheres of radius 10^-3 seen from a few units away (the reference's own
# discriminant is pure cancellation noise there: both must then be UNSURE or agree), and transforms at the eligibility
# limit (condition number 64).
REGIMES = {
    #            world scale, sphere radius range,   far end of the segment, condition range
    "millimetre": (1e-3, (0.15e-3, 2.5e-3), None, (1.0, 64.0)),
    "kilometre": (1e3, (0.15e3, 2.5e3), None, (1.0, 64.0)),
    "far light": (1.0, (0.15, 2.5), 1e4, (1.0, 64.0)),
    "tiny spheres": (1.0, (1e-3, 1e-2), None, (1.0, 64.0)),
    "huge spheres": (1.0, (1e2, 1e3), None, (1.0, 64.0)),
    "condition limit": (1.0, (0.15, 2.5), None, (40.0, 64.0)),
}


def regime_transforms(n, rng, world, radius, cond_range):
    def rot(axis, ang):
        c, s = np.cos(ang), np.sin(ang)
        r = np.tile(np.eye(3), (n, 1, 1))
        i, j = [(1, 2), (0, 2), (0, 1)][axis]
        r[:, i, i], r[:, j, j], r[:, i, j], r[:, j, i] = c, c, -s, s
        return r
    base = np.exp(rng.uniform(np.log(radius[0]), np.log(radius[1]), size=(n, 1)))
    aniso = np.exp(rng.uniform(0.0, np.log(cond_range[1]), size=(n, 3)))
    aniso[:, 0] = 1.0
    scale = base * aniso / aniso.max(axis=1, keepdims=True)  # the largest semi-axis is `base`
    fwd = rot(1, rng.uniform(0, 6.28, n)) @ rot(2, rng.uniform(0, 6.28, n)) @ (np.eye(3)[None] * scale[:, None, :])
    inv = np.linalg.inv(fwd)
    cond = np.abs(inv).sum(axis=2).max(axis=1) * np.abs(fwd).sum(axis=2).max(axis=1)
    keep = (cond <= cond_range[1]) & (cond >= cond_range[0])
    centre = rng.uniform(-4, 4, size=(n, 3)) * world
    t = -(inv @ centre[:, :, None])[:, :, 0]
    return inv[keep].astype(F), t[keep].astype(F), fwd[keep], centre[keep], cond[keep]


@pytest.mark.parametrize("regime", sorted(REGIMES))
def test_sphere_filter_at_scale_extremes(regime):
    world, radius, far, cond_range = REGIMES[regime]
    rng = np.random.default_rng(sorted(REGIMES).index(regime) + 40)
    inv, t, fwd, centre, cond = regime_transforms(500_000, rng, world, radius, cond_range)
    n = len(inv)
    assert n > 20_000, n
    tol = F(2.0 ** -24 * (128.0 * float(cond.max()) + 32.0))
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(u, rng.normal(size=(n, 3)))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    kind = rng.integers(0, 5, n)
    eps = 10.0 ** rng.uniform(-9, -1, n) * rng.choice([-1.0, 1.0], n)
    lift = np.where(kind == 0, 1.0 + eps, 1.0 + np.abs(rng.normal(0, 0.5, n)))
    a_obj = u * lift[:, None] - w * rng.uniform(0.5, 6, (n, 1))
    b_obj = u * lift[:, None] + w * rng.uniform(0.5, 6, (n, 1))
    a_obj = np.where((kind == 1)[:, None], u * (1.0 + eps)[:, None], a_obj)
    b_obj = np.where((kind == 2)[:, None], u * (1.0 + eps)[:, None], b_obj)
    p = (fwd @ a_obj[:, :, None])[:, :, 0] + centre
    light = (fwd @ b_obj[:, :, None])[:, :, 0] + centre
    rand = kind >= 3  # world-space ends a few world units away: for tiny spheres |w|^2 / R^2 reaches 10^7
    p[rand] = rng.normal(0, 2.5 * world, (rand.sum(), 3))
    light[rand] = rng.normal(0, 2.5 * world, (rand.sum(), 3))
    aim = rand & (rng.random(n) < 0.5)  # ... half of them aimed at the sphere so that hits occur at all
    light[aim] = (centre + (centre - p) * rng.uniform(0.2, 3.0, (n, 1)) + (fwd @ (u * rng.uniform(0, 1.3, (n, 1)))[:, :, None])[:, :, 0])[aim]
    if far is not None:
        d = light - p
        light = p + d / np.linalg.norm(d, axis=1, keepdims=True) * far * rng.uniform(0.3, 1.0, (n, 1))
    p, light = p.astype(F), light.astype(F)
    o = xf_point(inv, t, p)
    want = reference_sphere(inv, o, p, light)
    got = filter_sphere(inv, o, p, light, tol, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), (regime, int(wrong.sum()), np.flatnonzero(wrong)[:5], cond[wrong][:5])
    assert want[got == HIT].all() and 0.001 < (got == HIT).mean(), (regime, (got == HIT).mean())
    # decisiveness is a performance property, not a correctness one: every regime draws ellipsoids up to the 64:1
    # eligibility limit (bound 65 x 64 ulp) and 60 % of the segments are tangent / on-surface cases; a third is decided
    assert (got != UNSURE).mean() > 0.25, (regime, (got != UNSURE).mean())


@pytest.mark.parametrize("world,far", [(1e-3, None), (1e3, None), (1.0, 1e4), (1e3, 1e7)])
def test_plane_and_cube_filters_at_scale_extremes(world, far):
    rng = np.random.default_rng(int(abs(np.log10(world)) * 7 + (0 if far is None else 3)) + 50)
    n = 400_000
    # ---- planes: grazing segments, origin / light almost on the plane, at the world's scale
    normal = rng.normal(size=(n, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    normal *= np.exp(rng.uniform(np.log(0.1), np.log(10), (n, 1))) / world
    r1 = np.concatenate([normal, rng.uniform(-3, 3, (n, 1))], axis=1).astype(F)
    p = rng.normal(0, 3 * world, (n, 3))
    light = rng.normal(0, 3 * world, (n, 3))
    kind = rng.integers(0, 4, n)
    nn = normal / (np.linalg.norm(normal, axis=1, keepdims=True) ** 2)
    height = lambda q: (q * normal).sum(axis=1) + r1[:, 3]  # noqa: E731
    eps = 10.0 ** rng.uniform(-9, -2, n) * rng.choice([-1.0, 1.0], n)
    light = np.where((kind == 0)[:, None], light - nn * (height(light) - eps)[:, None], light)
    p = np.where((kind == 1)[:, None], p - nn * (height(p) - eps)[:, None], p)
    para = kind == 2
    light[para] = (light - nn * (height(light) - height(p) * (1.0 + eps * 50))[:, None])[para]
    if far is not None:
        d = light - p
        light = p + d / np.linalg.norm(d, axis=1, keepdims=True) * far * rng.uniform(0.3, 1.0, (n, 1))
    pf, lf = p.astype(F), light.astype(F)
    oy, want = reference_plane(r1, pf, lf)
    got = filter_plane(r1, oy, pf, lf, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), ("plane", int(wrong.sum()), np.flatnonzero(wrong)[:5])
    assert (got != UNSURE)[kind == 3].mean() > 0.98
    # ---- axis-aligned cubes (slabs down to 100:1) at the world's scale
    half = np.exp(rng.uniform(np.log(0.01), np.log(3), (n, 3))) * world
    centre = rng.uniform(-3, 3, (n, 3)) * world
    scale, t = (1.0 / half).astype(F), (-centre / half).astype(F)
    p = rng.normal(0, 3 * world, (n, 3))
    light = rng.normal(0, 3 * world, (n, 3))
    eps3 = 10.0 ** rng.uniform(-9, -2, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
    face = centre + half * np.where(rng.random((n, 3)) < 0.4, rng.choice([-1.0, 1.0], (n, 3)), rng.uniform(-1, 1, (n, 3)))
    light = np.where((kind == 0)[:, None], face + eps3 * half, light)
    p = np.where((kind == 1)[:, None], face + eps3 * half, p)
    graze = kind == 2
    light[graze] = (face + (face - p) * rng.uniform(0.1, 2, (n, 1)) + eps3 * half)[graze]
    if far is not None:
        d = light - p
        light = p + d / np.linalg.norm(d, axis=1, keepdims=True) * far * rng.uniform(0.3, 1.0, (n, 1))
    pf, lf = p.astype(F), light.astype(F)
    o, want = reference_cube(scale, t, pf, lf)
    got = filter_cube(scale, o, pf, lf, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), ("cube", int(wrong.sum()), np.flatnonzero(wrong)[:5])


@pytest.mark.parametrize("seed,world", [(61, 1.0), (62, 1e-3), (63, 1e3)])
def test_sphere_filter_with_diagonal_transforms_at_the_64_ulp_bound(seed, world):
    """Translation x axis-aligned scaling (every sphere of the reference's demo scenes): each direction component is one
    product, nothing cancels, and the commit ships tol = 64 u whatever the anisotropy (rtc_commit.cu: plan_small_scene;
    derivation: 48 u).  Ellipsoids up to 64:1, tangent / on-surface / random segments, three world scales."""
    rng = np.random.default_rng(seed)
    n = 500_000
    scale = np.exp(rng.uniform(np.log(0.05), np.log(3.2), size=(n, 3))) * world
    iso = rng.random(n) < 0.4
    scale[iso] = scale[iso, :1]
    fwd = np.eye(3)[None] * scale[:, None, :]
    inv = (np.eye(3)[None] / scale[:, None, :]).astype(F)
    centre = rng.uniform(-4, 4, size=(n, 3)) * world
    t = (-centre / scale).astype(F)
    tol = F(2.0 ** -24 * 64.0)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(u, rng.normal(size=(n, 3)))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    kind = rng.integers(0, 5, n)
    eps = 10.0 ** rng.uniform(-9, -1, n) * rng.choice([-1.0, 1.0], n)
    lift = np.where(kind == 0, 1.0 + eps, 1.0 + np.abs(rng.normal(0, 0.5, n)))
    a_obj = u * lift[:, None] - w * rng.uniform(0.5, 6, (n, 1))
    b_obj = u * lift[:, None] + w * rng.uniform(0.5, 6, (n, 1))
    a_obj = np.where((kind == 1)[:, None], u * (1.0 + eps)[:, None], a_obj)
    b_obj = np.where((kind == 2)[:, None], u * (1.0 + eps)[:, None], b_obj)
    rand = kind >= 3
    a_obj[rand] = rng.normal(0, 2.5, (rand.sum(), 3))
    b_obj[rand] = rng.normal(0, 2.5, (rand.sum(), 3))
    p = ((fwd @ a_obj[:, :, None])[:, :, 0] + centre).astype(F)
    light = ((fwd @ b_obj[:, :, None])[:, :, 0] + centre).astype(F)
    o = xf_point(inv, t, p)
    want = reference_sphere(inv, o, p, light)
    got = filter_sphere(inv, o, p, light, tol, rng)
    wrong = ((got == HIT) & ~want) | ((got == MISS) & want)
    assert not wrong.any(), (int(wrong.sum()), np.flatnonzero(wrong)[:5])
    assert (got != UNSURE)[rand].mean() > 0.97
