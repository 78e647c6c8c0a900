"""TEST INFRASTRUCTURE ONLY -- synthetic inputs for the MaD hot path.

Seeded random-walk "proteins" (3.8 A steps reflected inside a cubic box, all carbon), a PDB
text writer the reference's fixed-column parser accepts (``mad/PDB.py:41-65``), a NumPy-only
density simulator for benchmark-sized maps (used where ``/root/reference`` is absent, e.g. on
the GPU box), and the synthetic descriptor sets of config C5.

Nothing here is used by the product path; ``bench.py`` uses it only to *make inputs*.
"""
import math

import numpy as np


def random_walk_atoms(n_atoms, box, seed, step=3.8, origin=(0.0, 0.0, 0.0)):
    """Seeded 3-D random walk of ``n_atoms`` points with ``step`` A steps, reflected in [0, box]^3."""
    rng = np.random.default_rng(seed)
    pts = np.empty((n_atoms, 3), dtype=np.float64)
    p = rng.uniform(0.25 * box, 0.75 * box, size=3)
    for i in range(n_atoms):
        d = rng.normal(size=3)
        d *= step / math.sqrt(float(d @ d))
        p = p + d
        for a in range(3):
            if p[a] < 0.0:
                p[a] = -p[a]
            if p[a] > box:
                p[a] = 2.0 * box - p[a]
        pts[i] = p
    return pts + np.asarray(origin, dtype=np.float64)


def write_pdb(path, coords, chain="A"):
    """One CA/ALA/carbon ATOM record per point, in the column layout of ``mad/PDB.py:90``."""
    with open(path, "w") as f:
        for i, (x, y, z) in enumerate(coords):
            f.write("%-6s%5i  %-3s %3s%2s%4s    %8.3f%8.3f%8.3f%6.2f%6.2f          %-2s\n"
                    % ("ATOM", (i + 1) % 100000, "CA", "ALA", chain, (i + 1) % 10000,
                       x, y, z, 1.0, 0.0, "C"))


def quantise_u16(grid):
    """Map a [0,1] density to uint16 levels; ``dequantise_u16`` gives a portable, exact f32 grid."""
    g = np.asarray(grid, dtype=np.float64)
    return np.clip(np.rint(g * 65535.0), 0, 65535).astype(np.uint16)


def dequantise_u16(q):
    return (q.astype(np.float32) / np.float32(65535.0)).astype(np.float32)


def simulate_density(coords, resolution, voxelsp, margin=2):
    """NumPy-only density simulation in the spirit of ``mad/PDB.py:131-163`` (trilinear splat of
    unit masses + separable Gaussian of sigma = res/(pi*sqrt(2))/voxelsp truncated at 3 sigma,
    max-normalised).  NOT bit-identical to the reference's (it is only an input generator)."""
    coords = np.asarray(coords, dtype=np.float64)
    lo = voxelsp * np.floor(coords.min(0) / voxelsp)
    hi = voxelsp * np.ceil(coords.max(0) / voxelsp)
    sig = resolution / (math.pi * math.sqrt(2.0)) / voxelsp
    r = int(math.ceil(3.0 * sig))
    pad = margin + r
    dims = np.ceil((hi - lo) / voxelsp).astype(int) + 2 * pad + 1
    g = (coords - lo) / voxelsp + pad
    i0 = np.floor(g).astype(np.int64)
    t = g - i0
    grid = np.zeros(tuple(dims), dtype=np.float64)
    for dx in (0, 1):
        wx = t[:, 0] if dx else 1.0 - t[:, 0]
        for dy in (0, 1):
            wy = t[:, 1] if dy else 1.0 - t[:, 1]
            for dz in (0, 1):
                wz = t[:, 2] if dz else 1.0 - t[:, 2]
                np.add.at(grid, (i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz), wx * wy * wz)
    k = np.array([math.exp(-(j * j) / (2.0 * sig * sig)) for j in range(-r, r + 1)])
    k /= k.sum()
    for ax in range(3):
        acc = np.zeros_like(grid)
        n = grid.shape[ax]
        for j in range(-r, r + 1):
            src = [slice(None)] * 3
            dst = [slice(None)] * 3
            if j >= 0:
                src[ax] = slice(0, n - j)
                dst[ax] = slice(j, n)
            else:
                src[ax] = slice(-j, n)
                dst[ax] = slice(0, n + j)
            acc[tuple(dst)] += k[j + r] * grid[tuple(src)]
        grid = acc
    grid /= grid.max()
    origin = lo - pad * voxelsp
    return grid.astype(np.float32), origin


def fit_to_cube(grid, n):
    """Centre-crop / zero-pad ``grid`` to exactly n^3 (config C2/C3 ask for exact sizes)."""
    out = np.zeros((n, n, n), dtype=grid.dtype)
    src, dst = [], []
    for a in range(3):
        s = grid.shape[a]
        if s >= n:
            o = (s - n) // 2
            src.append(slice(o, o + n))
            dst.append(slice(0, n))
        else:
            o = (n - s) // 2
            src.append(slice(0, s))
            dst.append(slice(o, o + s))
    out[tuple(dst)] = grid[tuple(src)]
    return out


def assembly_atoms(n, voxelsp, n_sub, atoms_per_sub, seed0, box=None):
    """Atom sets of ``n_sub`` random-walk subunits packed on a coarse lattice inside an n^3 box of
    ``voxelsp`` A voxels (list of [atoms,3] arrays, one per subunit; seeds seed0 .. seed0+n_sub-1)."""
    side = n * voxelsp
    per_axis = int(math.ceil(n_sub ** (1.0 / 3.0)))
    cell = side * 0.8 / per_axis
    if box is None:
        box = cell * 0.92
    pts = []
    for s in range(n_sub):
        ix, iy, iz = s % per_axis, (s // per_axis) % per_axis, s // (per_axis * per_axis)
        org = (0.1 * side + ix * cell, 0.1 * side + iy * cell, 0.1 * side + iz * cell)
        pts.append(random_walk_atoms(atoms_per_sub, min(box, cell * 0.92), seed0 + s, origin=org))
    return pts


def assembly_map(n, resolution, voxelsp, n_sub, atoms_per_sub, seed0, box=None):
    """C2/C3-style map: the subunits of ``assembly_atoms`` simulated together at ``resolution`` A
    and fitted to exactly n^3."""
    pts = assembly_atoms(n, voxelsp, n_sub, atoms_per_sub, seed0, box)
    grid, _ = simulate_density(np.concatenate(pts), resolution, voxelsp)
    return fit_to_cube(grid, n)


def assembly_with_components(n, resolution, voxelsp, n_sub, atoms_per_sub, seed0, box=None):
    """(assembly map n^3, [component maps]) -- each component simulated alone from its own atoms
    (what MaD does with the subunit PDBs, mad/MapSpace.py:73-76), origin dropped."""
    pts = assembly_atoms(n, voxelsp, n_sub, atoms_per_sub, seed0, box)
    grid, _ = simulate_density(np.concatenate(pts), resolution, voxelsp)
    comps = [simulate_density(p, resolution, voxelsp)[0] for p in pts]
    return fit_to_cube(grid, n), comps


def representative_crop(grid, side, level=0.05, step=None):
    """Sub-cube of ``side``^3 whose occupancy (fraction of voxels > level) is closest to the whole
    map's: the bounded CPU sample of a benchmark-sized map (same voxels/s definition)."""
    n = grid.shape
    step = step or max(side // 2, 1)
    occ = grid > level
    target = float(occ.mean())
    best, best_err = (0, 0, 0), None
    for x in range(0, max(n[0] - side, 0) + 1, step):
        for y in range(0, max(n[1] - side, 0) + 1, step):
            for z in range(0, max(n[2] - side, 0) + 1, step):
                f = float(occ[x:x + side, y:y + side, z:z + side].mean())
                if best_err is None or abs(f - target) < best_err:
                    best, best_err = (x, y, z), abs(f - target)
    x, y, z = best
    return np.ascontiguousarray(grid[x:x + side, y:y + side, z:z + side]), best


def synthetic_descriptors(m, seed, noisy_copy_of=None, copy_frac=0.5, redraw=0.10):
    """C5 descriptor sets: int16[m,1024]; each 16-bin block ~ multinomial(64, Dirichlet(0.5)).
    With ``noisy_copy_of`` given, ``copy_frac`` of the rows are noisy copies of random rows of it
    (``redraw`` of the 64 votes of every block re-drawn), the rest fresh; rows are shuffled."""
    rng = np.random.default_rng(seed)

    def fresh(k):
        p = rng.dirichlet(np.full(16, 0.5), size=(k, 64))
        out = np.empty((k, 64, 16), dtype=np.int16)
        flat_p = p.reshape(-1, 16)
        out.reshape(-1, 16)[:] = rng.multinomial(64, flat_p)
        return out

    if noisy_copy_of is None:
        return fresh(m).reshape(m, 1024)
    n_copy = int(m * copy_frac)
    src = rng.integers(0, noisy_copy_of.shape[0], size=n_copy)
    base = noisy_copy_of[src].reshape(n_copy, 64, 16).astype(np.int64)
    n_re = int(round(64 * redraw))
    # remove n_re votes (proportionally to counts) and re-draw them uniformly
    for _ in range(n_re):
        cdf = np.cumsum(base, axis=-1)
        tot = cdf[..., -1:]
        u = rng.random(size=tot.shape) * tot
        pick = (cdf <= u).sum(-1)
        pick = np.minimum(pick, 15)
        has = tot[..., 0] > 0
        ii, jj = np.nonzero(has)
        base[ii, jj, pick[ii, jj]] -= 1
        add = rng.integers(0, 16, size=has.shape)
        base[ii, jj, add[ii, jj]] += 1
    out = np.concatenate([base.astype(np.int16).reshape(n_copy, 1024),
                          fresh(m - n_copy).reshape(m - n_copy, 1024)])
    perm = rng.permutation(m)
    return out[perm]


# ---------------------------------------------------------------------------------------------
# The benchmark workloads (BASELINE.json configs): ONE definition shared by bench.py (inputs only),
# the fixture generator (oracle/gen_golden_bench.py) and the parity tests, so that what is benchmarked is what is
# pinned against the reference.
# ---------------------------------------------------------------------------------------------
C2 = dict(n=256, resolution=8.0, voxelsp=2.0, n_sub=6, atoms_per_sub=40000, seed0=10, box=150.0)


def c2_inputs(rank=0):
    """(256^3 assembly map, [6 component maps]) of config C2; rank r > 0 gets its own map (seeds shifted by 100 r)."""
    cfg = dict(C2)
    cfg["seed0"] = C2["seed0"] + 100 * int(rank)
    return assembly_with_components(**cfg)


C4_MAPS = 64


def c4_snapshot(i, base=None):
    """Snapshot i of config C4: the C1 component (9 000-atom walk, seed 1) with every atom displaced by
    N(0, 1.5 A) (seed 100 + i), simulated at 4 A / 1 A per voxel and fitted to 96^3."""
    if base is None:
        base = random_walk_atoms(9000, 85.0, 1)
    rng = np.random.default_rng(100 + int(i))
    g, _ = simulate_density(base + rng.normal(scale=1.5, size=base.shape), 4.0, 1.0)
    return fit_to_cube(g, 96)


def c5_descriptor_sets(m, n):
    """(hi int16[m,1024], lo int16[n,1024]) of config C5.  8192 distinct multinomial/Dirichlet rows, tiled with a
    different cyclic shift per tile (keeps the host generation short); hi = 50 % noisy copies of lo rows."""
    base = synthetic_descriptors(8192, 7)
    reps = (n + len(base) - 1) // len(base)
    rng = np.random.default_rng(11)
    lo = np.concatenate([np.roll(base, int(rng.integers(0, 1024)), axis=1) if r else base for r in range(reps)])[:n]
    hi = synthetic_descriptors(min(m, 8192), 8, noisy_copy_of=lo[:8192])
    hi = np.concatenate([hi] * ((m + len(hi) - 1) // len(hi)))[:m]
    return np.ascontiguousarray(hi), np.ascontiguousarray(lo)
