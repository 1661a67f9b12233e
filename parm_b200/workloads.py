"""Synthetic inputs for the BASELINE.json configs.

Pure numpy, host side. Every generator returns a dict:
  ndim, L (ndim,), x (N,ndim), v (N,ndim), m (N,), kind, params (N,3), types (N,) uint32,
  eps_table (T,T), skin, dt, integrator, and for CollectionSol damping/T.
params columns follow the C ABI (include/parm_b200.h): (eps, sigma, exponent|sigcut).

Velocities: N(0,1)*sqrt(T/m), centre-of-mass velocity removed, rescaled so that the
reference's temp() formula (collection.cpp:135-142: 2*KE(v - v_com)/(ndof - NDIM))
gives exactly T.
"""
import numpy as np

KIND_LJREPULSE, KIND_REPULSION, KIND_LJATTRACTREPULSE, KIND_LJCUT = 0, 1, 2, 3
(KIND_LJATTRACTCUT, KIND_LJATTRACTFIXEDREPULSE, KIND_EISMCLACHLAN, KIND_LJISH, KIND_LJATTRACTREPULSESIGS,
 KIND_REPULSIONDRAG, KIND_LOISOHERN, KIND_LOISLIN, KIND_LOISOHERNMIN, KIND_LOISLINMIN) = range(4, 14)
VERLET, SOL = 0, 1


def tables(w):
    """(eps_table, sig_table) to hand to the interaction: the epsilon table only for the functors whose A struct
    carries `epsilons` (always for kinds 2, 5, 7; for kinds 1, 3, 4 when variant is "I" or "II"), the sigma table
    only for the doubly indexed A structs (variant "II": IEpsISigCutAtom / IEpsISigExpAtom)."""
    variant = str(w.get("variant", ""))
    indexed = int(w["kind"]) in (KIND_LJATTRACTREPULSE, KIND_LJATTRACTFIXEDREPULSE, KIND_LJISH) or variant in ("I", "II")
    return (w.get("eps_table") if indexed else None), (w.get("sig_table") if variant == "II" else None)


def _velocities(rng, n, ndim, T, m):
    v = rng.standard_normal((n, ndim)) * np.sqrt(T / m)[:, None]
    v -= (v * m[:, None]).sum(0) / m.sum()
    ke = 0.5 * (m[:, None] * v * v).sum()
    t = 2 * ke / (ndim * n - ndim)
    return v * np.sqrt(T / t)


def _lattice(shape, spacing):
    g = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij")
    return np.stack([q.ravel() for q in g], axis=1) * spacing


def lj_lattice(shape, rho=1.1939, T=1.44, jitter=0.05, skin=0.3, dt=0.004, sigcut=2.5, seed=3003,
               kind=KIND_LJATTRACTREPULSE):
    """Configs 3 and 5: one-species LJ (ParM sigma = potential minimum), simple-cubic start."""
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    n = int(np.prod(shape))
    a = rho ** (-1.0 / 3.0)
    x = _lattice(shape, a) + rng.uniform(-jitter, jitter, (n, 3)) * a + 0.5 * a
    L = np.array(shape, dtype=np.float64) * a
    m = np.ones(n)
    v = _velocities(rng, n, 3, T, m)
    params = np.zeros((n, 3))
    params[:, 0] = 1.0
    params[:, 1] = 1.0
    params[:, 2] = sigcut
    return dict(ndim=3, L=L, x=x, v=v, m=m, kind=kind, params=params, types=np.zeros(n, np.uint32),
                eps_table=np.ones((1, 1)), skin=skin, dt=dt, integrator=VERLET, name="lj3d_%dx%dx%d" % shape)


def config1(seed=1001, kind=KIND_LJCUT):
    """LJatoms.cpp-like: N=1000, phi=0.3, sigcut 2.5, skin 1.0, dt 1e-4 (LJatoms.cpp:11-17,26-27,38)."""
    rng = np.random.default_rng(seed)
    n = 1000
    L = (n * np.pi / 6 / 0.3) ** (1.0 / 3.0)
    a = L / 10
    x = _lattice((10, 10, 10), a) + rng.uniform(-0.1, 0.1, (n, 3)) * a + 0.5 * a
    m = np.ones(n)
    v = _velocities(rng, n, 3, 1.0, m)
    params = np.tile(np.array([1.0, 1.0, 2.5]), (n, 1))
    return dict(ndim=3, L=np.full(3, L), x=x, v=v, m=m, kind=kind, params=params, types=np.zeros(n, np.uint32),
                eps_table=np.ones((1, 1)), skin=1.0, dt=1e-4, integrator=VERLET, name="config1_lj1000")


def config2(nx=250, ny=400, seed=2002):
    """2-D bidisperse harmonic repulsion (RepulsionPair, exponent 2), phi=0.9 (tests.py:236-252)."""
    rng = np.random.default_rng(seed)
    n = nx * ny
    sig = np.where(rng.permutation(n) < n // 2, 1.0, 1.4)
    phi = 0.90
    area = (np.pi * sig ** 2 / 4).sum() / phi
    # rectangular cells so that the box is close to square
    ax = np.sqrt(area * ny / nx) / ny * 1.0
    Lx, Ly = None, None
    ax = np.sqrt(area / (nx * ny))
    Lx, Ly = nx * ax, ny * ax
    x = _lattice((nx, ny), ax) + rng.uniform(-0.2, 0.2, (n, 2)) * ax + 0.5 * ax
    m = np.ones(n)
    v = _velocities(rng, n, 2, 1e-3, m)
    params = np.stack([np.ones(n), sig, np.full(n, 2.0)], axis=1)
    return dict(ndim=2, L=np.array([Lx, Ly]), x=x, v=v, m=m, kind=KIND_REPULSION, params=params,
                types=np.zeros(n, np.uint32), eps_table=np.ones((1, 1)), skin=0.4, dt=0.01, integrator=VERLET,
                name="config2_harmonic2d_%d" % n)


def config3(side=100, **kw):
    return lj_lattice((side, side, side), seed=3003, **kw)


def config4(shape=(100, 200, 200), seed=4004):
    """3-D binary WCA-like LJRepulse glass former, CollectionSol xi=1, T=1."""
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    n = int(np.prod(shape))
    sig = np.where(rng.permutation(n) < n // 2, 1.0, 1.4)
    phi = 0.55
    rho = phi / (np.pi / 6 * (sig ** 3).mean())
    a = rho ** (-1.0 / 3.0)
    x = _lattice(shape, a) + rng.uniform(-0.05, 0.05, (n, 3)) * a + 0.5 * a
    L = np.array(shape, dtype=np.float64) * a
    m = np.ones(n)
    v = _velocities(rng, n, 3, 1.0, m)
    params = np.stack([np.ones(n), sig, np.zeros(n)], axis=1)
    return dict(ndim=3, L=L, x=x, v=v, m=m, kind=KIND_LJREPULSE, params=params, types=np.zeros(n, np.uint32),
                eps_table=np.ones((1, 1)), skin=0.3, dt=0.002, integrator=SOL, damping=1.0, T=1.0,
                name="config4_wca_%dx%dx%d" % shape)


def config5(shape=(200, 200, 400), **kw):
    return lj_lattice(shape, seed=5005, **kw)


def hertzian12(seed=131):
    """pyparm/tests.py:218-274 (RandomHertzianVerletTest): 3-D, N=12, Repulsion eps=1.2,
    sigma 1.0/1.4 half/half, exponent 2, m=sigma^3, phi=0.3, skin 0.4, dt 0.01; positions from
    numpy's legacy RandomState(seed) uniform in [0, L) exactly as tests.py:266-271 draws them."""
    n = 12
    sig = np.array([1.0] * (n // 2) + [1.4] * (n // 2))
    m = sig ** 3
    Vs = (np.pi / 6 * sig ** 3).sum()
    L = (Vs / 0.3) ** (1.0 / 3.0)
    rs = np.random.RandomState(seed)
    x = np.empty((n, 3))
    v = np.empty((n, 3))
    for i in range(n):  # tests.py:266-271 draws x, v, f per atom, interleaved
        x[i] = rs.uniform(0., L, size=(3,))
        v[i] = rs.normal(size=(3,)) / m[i]
        rs.normal(size=(3,))  # a.f, overwritten by set_forces
    params = np.stack([np.full(n, 1.2), sig, np.full(n, 2.0)], axis=1)
    return dict(ndim=3, L=np.full(3, L), x=x, v=v, m=m, kind=KIND_REPULSION, params=params,
                types=np.zeros(n, np.uint32), eps_table=np.ones((1, 1)), skin=0.4, dt=0.01, integrator=VERLET,
                name="hertzian12")


def random_system(n, ndim, kind, seed, rho=0.9, ntypes=2, polydisperse=True, T=1.0, skin=0.3, frozen=0,
                  continuous=False):
    """Small ragged random systems for parity tests: jittered lattice with vacancies, shifted by
    random whole box images, non-cubic box, per-atom eps/sigma/exponent/sigcut, several
    types with negative / zero table entries (LJAttractRepulsePair's repulsive-only and off branches)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n ** (1.0 / ndim)))
    a = rho ** (-1.0 / ndim)
    shape = [side] * ndim
    shape[-1] = side + 1 if ndim > 1 else side  # non-cubic box (box.hpp:102 allows per-axis sizes)
    L = np.array(shape, dtype=np.float64) * a
    sites = _lattice(shape, a)
    pick = rng.permutation(len(sites))[:n]      # ragged: random vacancies
    x = sites[pick] + rng.uniform(-0.12, 0.12, (n, ndim)) * a + 0.5 * a
    # unwrapped on purpose: ParM never wraps Atom::x, only diff() is periodic (trackers.cpp:27)
    x += rng.integers(-2, 3, (n, ndim)) * L
    m = rng.uniform(0.5, 2.0, n)
    if frozen:
        m[rng.choice(n, frozen, replace=False)] = 0.0
    mm = np.where(m > 0, m, 1.0)
    v = rng.standard_normal((n, ndim)) * np.sqrt(T / mm)[:, None]
    # discrete species (the device path tabulates pair constants per species pair)
    sig = rng.choice([0.8, 1.0, 1.2, 1.4], n) if polydisperse else np.ones(n)
    eps = rng.choice([0.5, 1.5], n)
    if continuous:  # every atom its own (eps, sigma): exercises the per-pair mixing path on the device
        sig = rng.uniform(0.8, 1.4, n)
        eps = rng.uniform(0.5, 1.5, n)
    third = {KIND_LJREPULSE: np.zeros(n), KIND_REPULSION: rng.choice([2.0, 2.5, 1.5], n),
             KIND_LJATTRACTREPULSE: rng.choice([2.0, 2.5], n), KIND_LJCUT: rng.choice([2.0, 2.5], n)}[kind]
    params = np.stack([eps, sig, third], axis=1)
    types = rng.integers(0, ntypes, n).astype(np.uint32)
    tab = rng.uniform(0.5, 1.5, (ntypes, ntypes))
    tab = (tab + tab.T) / 2
    if ntypes > 1:
        tab[0, 1] = tab[1, 0] = -0.7   # eps<=0 -> purely repulsive branch, interaction.hpp:1261-1266
    if ntypes > 2:
        tab[0, 2] = tab[2, 0] = 0.0    # eps==0 -> no force, interaction.hpp:1290
    return dict(ndim=ndim, L=L, x=x, v=v, m=m, kind=kind, params=params, types=types, eps_table=tab,
                skin=skin, dt=0.002, integrator=VERLET, name="random_%d_%dd_k%d" % (n, ndim, kind))


FUNCTOR_CASES = [  # (kind, variant): every NListed instantiation of sim.i:621-643 beyond the four hot-path ones
    (KIND_REPULSION, "II"), (KIND_LJCUT, "II"), (KIND_LJATTRACTCUT, ""), (KIND_LJATTRACTCUT, "I"),
    (KIND_LJATTRACTCUT, "II"), (KIND_LJATTRACTFIXEDREPULSE, ""), (KIND_EISMCLACHLAN, ""), (KIND_LJISH, ""),
    (KIND_LJATTRACTREPULSESIGS, ""), (KIND_REPULSIONDRAG, ""), (KIND_LOISOHERN, ""), (KIND_LOISLIN, ""),
    (KIND_LOISOHERNMIN, ""), (KIND_LOISLINMIN, "")]


def functor_system(kind, variant="", ndim=3, n=400, seed=0, continuous=False, T=0.2):
    """Ragged random system for one NListed functor. params columns follow
    include/parm_b200.h (parm_inter_set_params_ex). continuous=True gives every atom its own parameter
    tuple (> 32 species: the per-pair constructor then runs on the device)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n ** (1.0 / ndim)))
    a = 1.12
    shape = [side] * ndim
    shape[-1] = side + 1
    L = np.array(shape, dtype=np.float64) * a
    sites = _lattice(shape, a)
    pick = rng.permutation(len(sites))[:n]
    x = sites[pick] + rng.uniform(-0.13, 0.13, (n, ndim)) * a + 0.5 * a
    x += rng.integers(-1, 2, (n, ndim)) * L
    m = rng.uniform(0.5, 2.0, n)
    v = rng.standard_normal((n, ndim)) * np.sqrt(T / m)[:, None]
    nt = 3
    types = rng.integers(0, nt, n).astype(np.uint32)

    def pick2(lo, hi, choices):
        return rng.uniform(lo, hi, n) if continuous else rng.choice(choices, n)

    sig = pick2(0.9, 1.25, [1.0, 1.2])
    eps = pick2(0.5, 1.5, [0.6, 1.4])
    cut = rng.choice([1.8, 2.2], n)
    p = np.zeros((n, 5))
    p[:, 0], p[:, 1] = eps, sig
    tab = np.array([[1.0, -0.5, 0.0], [-0.5, 0.8, 0.3], [0.0, 0.3, 0.6]])
    stab = np.array([[1.0, 1.1, 1.2], [1.1, 1.25, 0.95], [1.2, 0.95, 1.05]])
    if kind == KIND_REPULSION:
        p[:, 2] = rng.choice([2.0, 2.5, 1.5], n)
        tab = np.abs(tab) + 0.2
    elif kind in (KIND_LJCUT, KIND_LJATTRACTCUT):
        p[:, 2] = cut
        if kind == KIND_LJCUT:
            stab = stab * 0.82  # full r^-12 core: keep sigma_ij below the lattice spacing or the trajectory is chaotic
        if kind == KIND_LJATTRACTCUT:
            tab = np.abs(tab)  # one zero entry: the eps == 0 early-out (interaction.hpp:214, :224)
    elif kind == KIND_LJATTRACTFIXEDREPULSE:
        p[:, 2], p[:, 3] = cut, pick2(0.8, 2.0, [1.0, 2.0])
    elif kind == KIND_EISMCLACHLAN:
        p[:, 0], p[:, 1] = pick2(-0.3, 0.4, [0.35, -0.2]), pick2(0.95, 1.35, [1.0, 1.3])
    elif kind == KIND_LJISH:
        p[:, 2], p[:, 3], p[:, 4] = cut, pick2(0.8, 2.0, [1.0, 2.0]), rng.choice([6.0, 4.0, 5.0], n)
    elif kind == KIND_LJATTRACTREPULSESIGS:
        p[:, 2], p[:, 3], p[:, 4] = cut, pick2(0.3, 0.9, [0.4, 0.8]), pick2(0.4, 0.7, [0.5, 0.6])
    elif kind == KIND_REPULSIONDRAG:
        p[:, 2], p[:, 3] = rng.choice([2.0, 2.5], n), pick2(0.1, 0.6, [0.2, 0.5])
    elif kind in (KIND_LOISOHERN, KIND_LOISOHERNMIN):
        p[:, 2], p[:, 3] = pick2(0.05, 0.25, [0.1, 0.2]), pick2(0.03, 0.15, [0.05, 0.1])
    elif kind in (KIND_LOISLIN, KIND_LOISLINMIN):
        width = pick2(0.1, 0.35, [0.2, 0.3])
        depth = pick2(0.02, 0.08, [0.03, 0.06])
        p[:, 2], p[:, 3] = depth / width, width
    w = dict(ndim=ndim, L=L, x=x, v=v, m=m, kind=kind, variant=variant, params=p, types=types, eps_table=tab,
             sig_table=stab, skin=0.3, dt=0.002, integrator=VERLET,
             name="functor_k%d%s_%dd%s" % (kind, variant, ndim, "_cont" if continuous else ""))
    return w


INTEGRATOR_CASES = {  # PARM_INTEG_* -> constructor arguments after dt
    2: ("damped", (0.8,)), 3: ("solht", (0.7, 0.4)), 4: ("overdamped", (0.5,)), 5: ("nosehoover", (3.0, 0.5)),
    6: ("gaussiant", ()), 7: ("gear3a", ()), 8: ("gear4a", (2,)), 9: ("gear5a", (2,)), 10: ("gear6a", (1,))}


def integrator_system(integ, ndim=3, n=150, seed=0, kind=None):
    """random_system + one of the (f)2 integrators; frozen atoms only where the integrator tests for them."""
    name, params = INTEGRATOR_CASES[integ]
    if kind is None:
        kind = KIND_LJATTRACTREPULSE if integ % 2 else KIND_REPULSION
    w = random_system(n, ndim, kind, seed=seed, ntypes=2, frozen=(2 if integ in (2, 3, 4) else 0), T=0.5)
    w.update(integrator=integ, integ_params=params, name="integ_%s_%dd" % (name, ndim))
    return w


def packer_system(ndim=3, n=200, seed=0, phi0=0.5, P0=1e-3):
    """pyparm/packmin.py:60-80: bidisperse (1 : 1.4) harmonic spheres dropped at random in a box of packing
    fraction phi0, masses sigma^ndim, NeighborList skin 0.4, CollectionNLCG(dt=0.1, P0, kappa=10, kmax=1000,
    secmax=40, seceps=1e-20)."""
    rng = np.random.default_rng(seed)
    sig = np.where(np.arange(n) < n // 2, 1.0, 1.4)
    Vs = (sig ** ndim).sum() * np.pi / (2 * ndim)
    L = (Vs / phi0) ** (1.0 / ndim)
    x = rng.random((n, ndim)) * L
    p = np.stack([np.ones(n), sig, np.full(n, 2.0)], axis=1)
    return dict(ndim=ndim, L=np.full(ndim, L), x=x, v=np.zeros((n, ndim)), m=sig ** ndim, kind=KIND_REPULSION, params=p,
                types=np.zeros(n, np.uint32), eps_table=np.ones((1, 1)), skin=0.4, dt=0.1, integrator=11,
                integ_params=(P0, 10.0, 1000, 40, 1e-20), name="packer_%dd_%d" % (ndim, n))
