"""Initial conditions of the reference's six samples (src/sample/*.cpp), as numpy structured
arrays with the SPHParticle layout.  Same lattices, same accumulation order for the running
coordinates (``x += dx`` in the reference is a sequential sum, reproduced with cumsum), so the
particle sets are the ones `./sph <sample>` would build for the same N.

Only members the reference's generators set are set here; everything else is zero
(the reference leaves them uninitialised; Solver::initialize sets alpha/balsara/sound).
"""
import numpy as np


def particle_dtype(dim):
    """In-memory layout of sph::SPHParticle (include/particle.hpp:8-33): 144/176/208 bytes."""
    v = (np.float64, (dim,))
    return np.dtype([
        ("pos", *v), ("vel", *v), ("vel_p", *v), ("acc", *v),
        ("mass", "f8"), ("dens", "f8"), ("pres", "f8"), ("ene", "f8"), ("ene_p", "f8"),
        ("dene", "f8"), ("sml", "f8"), ("sound", "f8"), ("balsara", "f8"), ("alpha", "f8"),
        ("gradh", "f8"), ("phi", "f8"), ("id", "i4"), ("neighbor", "i4"), ("next", "u8"),
    ], align=True)


def _alloc(n, dim):
    p = np.zeros(n, dtype=particle_dtype(dim))
    p["id"] = np.arange(n, dtype=np.int32)
    return p


def _running(x0, dx, count):
    """x0, x0+dx, (x0+dx)+dx, ... : the reference's `x += dx` running sum."""
    a = np.full(count, dx, dtype=np.float64)
    a[0] = x0
    return np.cumsum(a)


def shock_tube(N, gamma):
    """src/sample/shock_tube.cpp:18-51 (DIM=1)."""
    dx_r = 0.5 / N
    dx_l = dx_r * 0.25
    num = N * 10
    p = _alloc(num, 1)
    x, dx, dens, pres = -0.5 + dx_l * 0.5, dx_l, 1.0, 1.0
    mass = 0.5 / N * 0.25
    left = True
    for i in range(num):
        p["pos"][i, 0] = x
        p["dens"][i] = dens
        p["pres"][i] = pres
        x += dx
        if x > 0.5 and left:
            x, dx, dens, pres, left = 0.5 + dx_r * 0.5, dx_r, 0.25, 0.1795, False
    p["mass"] = mass
    p["ene"] = p["pres"] / ((gamma - 1.0) * p["dens"])
    return p


def _row(x0, step, xmax, cache):
    """One lattice row: x = x0, x0+step, ... emitted while x <= xmax (the reference emits a
    particle at x, advances, and ends the row once x > xmax)."""
    key = (x0, step)
    if key not in cache:
        cnt = int((xmax - x0) / step) + 3
        r = _running(x0, step, cnt)
        cache[key] = r[:int(np.argmax(r > xmax))]
    return cache[key]


def khi(N, gamma):
    """src/sample/khi.cpp:18-73 (DIM=2)."""
    num = N * N * 3 // 4
    dx = 1.0 / N
    mass = 1.5 / num
    xs_l, ys_l, reg_l, cache = [], [], [], {}
    y = dx * 0.5
    region, odd = 1, True
    x0 = dx * 0.5
    total = 0
    while total < num:
        r = _row(x0, 2.0 * dx if region == 1 else dx, 1.0, cache)
        xs_l.append(r)
        ys_l.append(np.full(len(r), y))
        reg_l.append(np.full(len(r), region))
        total += len(r)
        # end of row (x > 1.0): src/sample/khi.cpp:47-69
        y += dx
        region = 2 if (y > 0.25 and y < 0.75) else 1
        if region == 1:
            if odd:
                odd, x0 = False, dx * 1.5
            else:
                odd, x0 = True, dx * 0.5
        else:
            x0 = dx * 0.5
    xs = np.concatenate(xs_l)[:num]
    yy = np.concatenate(ys_l)[:num]
    reg = np.concatenate(reg_l)[:num]
    p = _alloc(num, 2)
    p["pos"][:, 0] = xs
    p["pos"][:, 1] = yy
    p["vel"][:, 0] = np.where(reg == 1, -0.5, 0.5)
    sigma2_inv = 2 / (0.05 * 0.05)
    p["vel"][:, 1] = 0.1 * np.sin(4.0 * np.pi * xs) * (
        np.exp(-(yy - 0.25) ** 2 * 0.5 * sigma2_inv) + np.exp(-(yy - 0.75) ** 2 * 0.5 * sigma2_inv))
    p["mass"] = mass
    p["dens"] = reg.astype(np.float64)
    p["pres"] = 2.5
    p["ene"] = p["pres"] / ((gamma - 1.0) * p["dens"])
    return p


def _square_lattice(N):
    """x,y of the N*N lattice on [-0.5,0.5]^2 with running sums (gresho/pairing generators)."""
    dx = 1.0 / N
    row = _row(-0.5 + dx * 0.5, dx, 0.5, {})
    nrow = -(-N * N // len(row))
    ycol = _running(-0.5 + dx * 0.5, dx, nrow)
    xs = np.tile(row, nrow)[:N * N]
    ys = np.repeat(ycol, len(row))[:N * N]
    return xs, ys


def gresho_chan_vortex(N, gamma):
    """src/sample/gresho_chan_vortex.cpp:14-76 (DIM=2)."""
    num = N * N
    xs, ys = _square_lattice(N)
    p = _alloc(num, 2)
    p["pos"][:, 0], p["pos"][:, 1] = xs, ys
    r = np.sqrt(xs * xs + ys * ys)
    vel = np.where(r < 0.2, 5.0 * r, np.where(r < 0.4, 2.0 - 5.0 * r, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        pres = np.where(r < 0.2, 5.0 + 12.5 * r * r,
                        np.where(r < 0.4, 9.0 + 12.5 * r * r - 20.0 * r + 4.0 * np.log(5.0 * r),
                                 3.0 + 4.0 * np.log(2.0)))
        p["vel"][:, 0] = (-ys / r) * vel
        p["vel"][:, 1] = (xs / r) * vel
    p["dens"] = 1.0
    p["pres"] = pres
    p["mass"] = 1.0 / num
    p["ene"] = p["pres"] / ((gamma - 1.0) * p["dens"])
    return p


def pairing_instability(N, gamma):
    """src/sample/pairing_instability.cpp:16-55 (DIM=2): lattice + mt19937(1) jitter.

    numpy's MT19937 seeded through the legacy interface reproduces std::mt19937(1)'s 32-bit
    stream; std::uniform_real_distribution<double> (libstdc++ generate_canonical) consumes two
    32-bit words per draw: (lo + hi * 2^32) / 2^64.
    """
    num = N * N
    dx = 1.0 / N
    xs, ys = _square_lattice(N)
    rs = np.random.RandomState(1)
    w = rs.randint(0, 2 ** 32, size=4 * num, dtype=np.uint64).astype(np.float64)
    canon = (w[0::2] + w[1::2] * 4294967296.0) / 18446744073709551616.0
    canon = np.where(canon >= 1.0, np.nextafter(1.0, 0.0), canon)
    a, b = -dx * 0.05, dx * 0.05
    jit = canon * (b - a) + a
    p = _alloc(num, 2)
    p["pos"][:, 0] = xs + jit[0::2]
    p["pos"][:, 1] = ys + jit[1::2]
    p["dens"] = 1.0
    p["pres"] = 1.0
    p["mass"] = 1.0 / num
    p["ene"] = p["pres"] / ((gamma - 1.0) * p["dens"])
    return p


def hydrostatic(N, gamma):
    """src/sample/hydrostatic.cpp:14-76 (DIM=2)."""
    dx1 = 0.5 / N
    dx2 = dx1 * 2.0
    mass = 1.0 / (N * N)
    pts = []
    x = -0.25 + dx1 * 0.5
    y = -0.25 + dx1 * 0.5
    while y < 0.25:
        pts.append((x, y, 4.0))
        x += dx1
        if x > 0.25:
            x = -0.25 + dx1 * 0.5
            y += dx1
    x = -0.5 + dx2 * 0.5
    y = -0.5 + dx2 * 0.5
    while y < 0.5:
        pts.append((x, y, 1.0))
        while True:
            x += dx2
            if x > 0.5:
                x = -0.5 + dx2 * 0.5
                y += dx2
            if not (-0.25 < x < 0.25 and -0.25 < y < 0.25):
                break
    a = np.array(pts)
    p = _alloc(len(a), 2)
    p["pos"][:, 0], p["pos"][:, 1] = a[:, 0], a[:, 1]
    p["mass"] = mass
    p["dens"] = a[:, 2]
    p["pres"] = 2.5
    p["ene"] = p["pres"] / ((gamma - 1.0) * p["dens"])
    return p


def evrard(N, gamma, G=1.0):
    """src/sample/evrard.cpp:19-63 (DIM=3): N^3 lattice on [-1,1]^3 clipped to r<=1, r -> r^1.5."""
    dx = 2.0 / N
    c = (np.arange(N) + 0.5) * dx - 1.0
    # i outermost, k innermost (evrard.cpp:23-25)
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    X, Y, Z = X.ravel(), Y.ravel(), Z.ravel()
    r0 = np.sqrt(X * X + Y * Y + Z * Z)
    keep = ~(r0 > 1.0)
    X, Y, Z, r0 = X[keep], Y[keep], Z[keep], r0[keep]
    pos = r0 > 0.0
    scale = np.ones_like(r0)
    scale[pos] = np.power(r0[pos], 1.5) / r0[pos]
    n = len(X)
    p = _alloc(n, 3)
    p["pos"][:, 0], p["pos"][:, 1], p["pos"][:, 2] = X * scale, Y * scale, Z * scale
    u = 0.05 * G
    p["mass"] = 1.0 / n
    q = p["pos"]
    with np.errstate(divide="ignore"):
        p["dens"] = 1.0 / (2.0 * np.pi * np.sqrt(q[:, 0] ** 2 + q[:, 1] ** 2 + q[:, 2] ** 2))
    p["ene"] = u
    p["pres"] = (gamma - 1.0) * p["dens"] * u
    return p


def make_sample(params):
    """Solver::make_initial_condition (src/solver.cpp:476-495) for resolved sample params."""
    name, N, gamma = params["sample"], params["N"], params["gamma"]
    if name == "shock_tube":
        return shock_tube(N, gamma)
    if name == "khi":
        return khi(N, gamma)
    if name == "gresho_chan_vortex":
        return gresho_chan_vortex(N, gamma)
    if name == "pairing_instability":
        return pairing_instability(N, gamma)
    if name == "hydrostatic":
        return hydrostatic(N, gamma)
    if name == "evrard":
        return evrard(N, gamma, params["G"])
    raise ValueError("unknown sample type.")
