"""Synthetic workloads of bench.py and the full-size tests (host side, numpy only; nothing here computes sigma).

The generators follow SURVEY.md 8(d): the BASELINE configs whose real integrals cannot travel (no network, no psi4) are
replaced by seeded synthetic integrals of the same shape and symmetry:  chemist (ik|jl) = sum_P B^P_ik B^P_jl  with every
B^P symmetric and belonging to one irrep, so the table has the 8-fold permutational symmetry and the point-group
selection rules the reference's Hamiltonian class asserts (Hamiltonian.cpp:108-127).
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# N2 cc-pVDZ (reference tests/matrixelements/N2.CCPVDZ.FCIDUMP header: 28 orbitals, 14 electrons, D2h): orbital irreps in
# psi4 numbering AFTER Problem::SetupReorderD2h (Problem.cpp:57-95): Ag x7, B1u x7, B3u x3, B2g x3, B2u x3, B3g x3, B1g, Au
N2_CCPVDZ_IRREPS = [0] * 7 + [5] * 7 + [7] * 3 + [2] * 3 + [6] * 3 + [3] * 3 + [1] + [4]


def synthetic_integrals(L, irreps, seed, naux, width, amp, local=False):
    """-> (tmat[L,L], vmat[L,L,L,L] physicist <ab|cd> = (ac|bd)) ; seeded, symmetric, irrep-adapted"""
    rng = np.random.default_rng(seed)
    irr = np.asarray(irreps)
    nirr_prod = irr[:, None] ^ irr[None, :]
    idx = np.arange(L)
    B = np.zeros((naux, L, L))
    groups = sorted(set(nirr_prod.ravel().tolist()))
    for P in range(naux):
        c = rng.uniform(0, L)
        g = groups[P % len(groups)] if len(groups) > 1 else 0
        xi = rng.standard_normal((L, L))
        xi = 0.5 * (xi + xi.T)
        m = amp * np.exp(-(np.abs(idx[:, None] - c) + np.abs(idx[None, :] - c)) / width) * (1.0 + 0.1 * xi)
        if local:
            m = m * np.exp(-np.abs(idx[:, None] - idx[None, :]) / 1.5)
        B[P] = np.where(nirr_prod == g, m, 0.0)
    chem = np.einsum("pik,pjl->ikjl", B, B, optimize=True)          # (ik|jl)
    vmat = np.ascontiguousarray(chem.transpose(0, 2, 1, 3))         # <ij|kl> = (ik|jl)
    xi = rng.standard_normal((L, L))
    xi = 0.5 * (xi + xi.T)
    t = -0.5 * np.exp(-np.abs(idx[:, None] - idx[None, :]) / 1.5) * (1.0 + 0.1 * xi)
    t[idx, idx] = -1.0 + 0.05 * rng.standard_normal(L)
    tmat = np.where(nirr_prod == 0, t, 0.0)
    return tmat, vmat


# carbon positions (Angstrom) of tetracene, reference sphinx/tetracene.fcidump.in:6-23
TETRACENE_C = [(4.888883611380, -0.715374463486), (4.888883611380, 0.715374463486), (-4.888883611380, -0.715374463486),
               (-4.888883611380, 0.715374463486), (3.711144499602, -1.409316610825), (3.711144499602, 1.409316610825),
               (-3.711144499602, -1.409316610825), (-3.711144499602, 1.409316610825), (2.450542389320, -0.725895641808),
               (2.450542389320, 0.725895641808), (-2.450542389320, -0.725895641808), (-2.450542389320, 0.725895641808),
               (1.235393613403, -1.406341384439), (1.235393613403, 1.406341384439), (-1.235393613403, -1.406341384439),
               (-1.235393613403, 1.406341384439), (0.000000000000, -0.726150477978), (0.000000000000, 0.726150477978)]


def ppp_tetracene_integrals():
    """SURVEY.md 8(d) config 3 stand-in (the 20 GB psi4 FCIDUMP of sphinx/handson.rst:57 cannot be regenerated here): 18e/18o
    Pariser-Parr-Pople model on the tetracene carbon skeleton.  t = -2.4 eV for C-C < 1.6 A, Ohno gamma_ij = U / sqrt(1 + (U r_ij /
    14.397)^2) with U = 11.26 eV, <ij|kl> = delta_ik delta_jl gamma_ij, T_ii = -sum_{j != i} gamma_ij, everything / 27.2114 (Hartree);
    sites ordered along the long molecular axis (x, then y).  Known answer from the unmodified reference:
    E(D = 600) = -23.892594067 Eh.  -> (tmat[L,L], vmat[L,L,L,L] physicist)"""
    xy = np.array(sorted(TETRACENE_C, key=lambda p: (round(p[0], 6), round(p[1], 6))))
    L = len(xy)
    r = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    U = 11.26
    gamma = U / np.sqrt(1.0 + (U * r / 14.397) ** 2)
    t = np.zeros((L, L))
    t[(r < 1.6) & (r > 1e-9)] = -2.4
    for i in range(L):
        t[i, i] = -(gamma[i].sum() - gamma[i, i])
    v = np.zeros((L, L, L, L))
    for i in range(L):
        for j in range(L):
            v[i, j, i, j] = gamma[i, j]
    return t / 27.2114, v / 27.2114


class Workload:
    def __init__(self, name, L, group, N, twoS, irrep, irreps, D, site, seed, **gen):
        self.name, self.L, self.group, self.N, self.twoS, self.irrep = name, L, group, N, twoS, irrep
        self.irreps, self.D, self.site, self.seed, self.gen = list(irreps), D, site, seed, gen
        self._ints = None

    def integrals(self):
        if self._ints is None:
            self._ints = ppp_tetracene_integrals() if self.gen.get("model") == "ppp_tetracene" else synthetic_integrals(self.L, self.irreps, self.seed, **self.gen)
        return self._ints

    def context(self, device):
        from . import api
        t, v = self.integrals()
        ctx = api.Context(device)
        # column-major (first index fastest) flat tables, as b2_problem_set_integrals expects
        ctx.set_problem(self.L, self.group, self.N, self.twoS, self.irrep, self.irreps, tmat=t.ravel(order="F"), vmat=v.ravel(order="F"))
        ctx.bk_init(self.D)
        return ctx

    def sector_rows(self, ctx):
        """[(boundary, N, twoS, irrep, fcidim)] for every sector of the bookkeeper"""
        from ._lib import lib
        rows = []
        nirr = {0: 1, 1: 2, 2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 7: 8}[self.group]
        for b in range(self.L + 1):
            for n in range(lib.b2_bk_nmin(ctx.h, b), lib.b2_bk_nmax(ctx.h, b) + 1):
                for ts in range(lib.b2_bk_twosmin(ctx.h, b, n), lib.b2_bk_twosmax(ctx.h, b, n) + 1, 2):
                    for ir in range(nirr):
                        rows.append((b, n, ts, ir, ctx.fcidim(b, n, ts, ir)))
        return rows

    def apply_distribution(self, ctx, dist="flat", sigma_n=1.6, sigma_s=1.3):
        """Virtual-dimension distribution over the symmetry sectors of every boundary.
        'flat'  = SyBookkeeper's initial FCI-scaled distribution (what the reference starts its first sweep with);
        'gauss' = model of a converged state: dims ~ D * exp(-(N-Nbar)^2/2sn^2 - S^2/2ss^2), capped by the FCI dims
                  (O(50) populated sectors with leading dims of 0.05-0.1 D).
        -> int32 array (n,5) of (boundary, N, twoS, irrep, dim) actually set"""
        rows = self.sector_rows(ctx)
        out = []
        if dist == "flat":
            for b, n, ts, ir, fci in rows:
                out.append((b, n, ts, ir, ctx.dim(b, n, ts, ir)))
            return np.array(out, dtype=np.int32)
        assert dist == "gauss", dist
        by_b = {}
        for r in rows:
            by_b.setdefault(r[0], []).append(r)
        for b, lst in by_b.items():
            nbar = self.N * b / self.L
            fsum = {}
            for _, n, ts, ir, fci in lst:
                fsum[(n, ts)] = fsum.get((n, ts), 0) + fci
            raw = np.array([np.exp(-(n - nbar) ** 2 / (2 * sigma_n ** 2) - (ts / 2.0) ** 2 / (2 * sigma_s ** 2)) * (fci / fsum[(n, ts)] if fci else 0.0)
                            for _, n, ts, ir, fci in lst])
            tot = raw.sum()
            for (bb, n, ts, ir, fci), w in zip(lst, raw):
                d = int(min(fci, np.floor(self.D * w / tot + 0.5))) if tot > 0 else 0
                if b == 0 or b == self.L:
                    d = min(fci, 1)
                out.append((bb, n, ts, ir, d))
        arr = np.array(out, dtype=np.int32)
        ctx.bk_import(np.concatenate([arr, np.zeros((len(arr), 1), dtype=np.int32)], axis=1))
        return arr

    def write_problem_file(self, path):
        """binary problem file read by `ref_driver synth`"""
        t, v = self.integrals()
        with open(path, "wb") as f:
            np.array([self.L, self.group, self.N, self.twoS, self.irrep], dtype="<i4").tofile(f)
            np.array(self.irreps, dtype="<i4").tofile(f)
            np.array([0.0], dtype="<f8").tofile(f)
            t.ravel(order="F").astype("<f8").tofile(f)
            v.ravel(order="F").astype("<f8").tofile(f)

    def describe(self):
        return f"{self.name}: {self.N}e/{self.L}o group {self.group} D={self.D} site pair ({self.site},{self.site + 1})"


class FixtureWorkload(Workload):
    """a real Hamiltonian whose folded integral table travels as a problem-only fixture (tests/golden/problem_*.npz, written by
    tests/golden/make_golden.py from the reference's own FCIDUMP through the reference's Hamiltonian / Problem classes)"""

    def __init__(self, name, fixture, D, site):
        fx = np.load(os.path.join(ROOT, "tests", "golden", fixture))
        L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]
        super().__init__(name, L, group, N, twoS, irrep, [int(x) for x in fx["problem/orb_irrep"]], D, site, 0)
        self._mx, self._econst = fx["problem/mx"], float(fx["problem/econst"][0])

    def context(self, device):
        from . import api
        ctx = api.Context(device)
        ctx.set_problem(self.L, self.group, self.N, self.twoS, self.irrep, self.irreps, mx=self._mx, econst=self._econst)
        ctx.bk_init(self.D)
        return ctx


def get(name, D=None, site=None):
    """named workloads = the BASELINE.json configs (shape + symmetry), synthetic integrals"""
    if name == "synth40":      # config 5: 40e/40o C1, no locality
        w = Workload(name, 40, 0, 40, 0, 0, [0] * 40, 4000, 19, 40404000, naux=80, width=6.0, amp=0.25)
    elif name == "synth60":    # config 4: 60e/60o 1-D chain
        w = Workload(name, 60, 0, 60, 0, 0, [0] * 60, 4000, 29, 60604000, naux=120, width=2.0, amp=0.6, local=True)
    elif name == "n2":         # config 2: N2/cc-pVDZ shape (14e/28o, D2h, reordered)
        w = Workload(name, 28, 7, 14, 0, 0, N2_CCPVDZ_IRREPS, 2000, 13, 14282000, naux=96, width=6.0, amp=0.3)
    elif name == "tetracene":  # config 3: 18e/18o C1
        w = Workload(name, 18, 0, 18, 0, 0, [0] * 18, 3000, 8, 18183000, naux=36, width=3.0, amp=0.4, local=True)
    elif name == "n2_ccpvdz":   # config 2 itself: N2/cc-pVDZ 14e/28o D2h X1Sigma_g+, orbitals reordered by Problem::SetupReorderD2h;
        # published (sphinx/resources.rst:50-56): E(D=1000) = -109.28209711, E(D=2000) = -109.28216077 (219 s/sweep on 16 cores)
        w = FixtureWorkload(name, "problem_n2_ccpvdz.npz", 2000, 13)
    elif name == "tetracene_ppp":   # config 3 stand-in with a known answer from the reference
        w = Workload(name, 18, 0, 18, 0, 0, [0] * 18, 600, 8, 0, model="ppp_tetracene")
    elif name == "tiny":       # CPU-checkable stand-in used by smoke() and the fast tests
        w = Workload(name, 8, 5, 8, 0, 0, [0, 0, 2, 3, 0, 0, 2, 3], 40, 3, 8080, naux=16, width=3.0, amp=0.4)
    else:
        raise KeyError(name)
    if D is not None:
        w.D = int(D)
    if site is not None:
        w.site = int(site)
    return w
