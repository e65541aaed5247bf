"""The C++ mirror of the reference's public classes (include/chemps2_b200.hpp) driven by a caller written like the reference's own
tests (tests/cpp/dmrg_caller.cpp ~ tests/test1/2/5.cpp.in).  CPU part: Hamiltonian (FCIDUMP reader / writer), Problem (orbital
reordering, folded integral table) against the tables the unmodified reference dumped into tests/golden/.  GPU part: a whole
calculation (Solve, calc2DMandCorrelations, excited states) against the known answers of the reference's test5."""
import json
import os
import subprocess

import numpy as np
import pytest

from chemps2_b200 import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALLER = os.path.join(ROOT, "tests", "cpp", "_bin", "dmrg_caller")
GOLDEN = os.path.join(ROOT, "tests", "golden")
PSI2MOLPRO = {0: [1], 5: [1, 4, 2, 3], 7: [1, 4, 6, 7, 8, 5, 3, 2]}     # Irreps.cpp:184-216


def write_fcidump(path, L, group, N, twoS, irrep, orb_irrep, tmat, vmat, econst):
    """FCIDUMP (chemist (ij|kl), 1-based, molpro ORBSYM) from T[i + L j] and physicist V[a + L(b + L(c + L d))] = <ab|cd>"""
    p2m = PSI2MOLPRO[group]
    T = np.asarray(tmat).reshape(L, L, order="F")
    V = np.asarray(vmat).reshape(L, L, L, L, order="F")
    with open(path, "w") as f:
        f.write(f" &FCI NORB= {L},NELEC= {N},MS2= {twoS},\n  ORBSYM=" + "".join(f"{p2m[int(i)]}," for i in orb_irrep) + f"\n  ISYM={p2m[irrep]},\n /\n")
        for i in range(L):
            for j in range(i + 1):
                for k in range(i + 1):
                    for l in range(k + 1):
                        v = V[i, k, j, l]
                        if v != 0.0:
                            f.write(f" {v:23.16E} {i + 1:3d} {j + 1:3d} {k + 1:3d} {l + 1:3d}\n")
        for i in range(L):
            for j in range(i + 1):
                if T[i, j] != 0.0:
                    f.write(f" {T[i, j]:23.16E} {i + 1:3d} {j + 1:3d}   0   0\n")
        f.write(f" {econst:23.16E}   0   0   0   0\n")


def fcidump_from_fixture(fx, path, N=None, twoS=None, irrep=None):
    L, group, n0, s0, i0 = [int(x) for x in fx["problem/hdr"]]
    N, twoS, irrep = (n0 if N is None else N), (s0 if twoS is None else twoS), (i0 if irrep is None else irrep)
    write_fcidump(path, L, group, N, twoS, irrep, fx["problem/orb_irrep"], fx["problem/tmat"], fx["problem/vmat"], float(fx["problem/econst"][0]))
    return L, group, N, twoS, irrep


def run_host(tmp_path, fx, reorder, **target):
    dump = str(tmp_path / "in.fcidump")
    L, group, N, twoS, irrep = fcidump_from_fixture(fx, dump, **target)
    out = str(tmp_path / "host.bin")
    res = subprocess.run([CALLER, "host", dump, str(group), str(twoS), str(N), str(irrep), reorder, out], check=True, capture_output=True, text=True)
    info = json.loads(res.stdout.strip().splitlines()[-1])
    raw = open(out, "rb").read()
    assert int(np.frombuffer(raw[:4], dtype=np.int32)[0]) == L
    irreps = np.frombuffer(raw[4:4 + 4 * L], dtype=np.int32)
    econst = float(np.frombuffer(raw[4 + 4 * L:12 + 4 * L], dtype=np.float64)[0])
    mx = np.frombuffer(raw[12 + 4 * L:], dtype=np.float64)
    assert mx.size == L ** 4
    return info, irreps, econst, mx


def test_caller_is_built():
    assert os.path.exists(CALLER), "make builds tests/cpp/_bin/dmrg_caller"


@pytest.mark.parametrize("name", ["n2_sto3g_singlet", "h2o_631g", "hubbard10_sextet"])
def test_hamiltonian_problem_fold_matches_reference_table(tmp_path, name):
    """FCIDUMP -> Hamiltonian -> Problem::construct_mxelem equals the gMxElement table of the unmodified reference (Problem.cpp:363-384)"""
    fx = fixtures.load(os.path.join(GOLDEN, name + ".npz"))
    info, irreps, econst, mx = run_host(tmp_path, fx, "none")
    assert info["reorder"] == 0 and info["fcidump_roundtrip"] == 0.0
    assert np.array_equal(irreps, fx["problem/orb_irrep"])
    assert econst == float(fx["problem/econst"][0])
    assert np.abs(mx - fx["problem/mx"]).max() < 1e-15


def test_reorder_d2h_matches_reference_table(tmp_path):
    """Problem::SetupReorderD2h (Problem.cpp:57-94): the quintet B1u fixture was dumped by the reference WITH the reordering from the same
    N2/STO-3G integrals the singlet fixture holds in Hamiltonian order"""
    ham = fixtures.load(os.path.join(GOLDEN, "n2_sto3g_singlet.npz"))
    dmrg = fixtures.load(os.path.join(GOLDEN, "n2_sto3g_quintet_b1u.npz"))
    _, _, N, twoS, irrep = [int(x) for x in dmrg["problem/hdr"]]
    info, irreps, econst, mx = run_host(tmp_path, ham, "d2h", N=N, twoS=twoS, irrep=irrep)
    assert info["reorder"] == 1
    assert np.array_equal(irreps, dmrg["problem/orb_irrep"])
    assert np.abs(mx - dmrg["problem/mx"]).max() < 1e-15


def test_reorder_c2v_restated(tmp_path):
    """Problem::SetupReorderC2v (Problem.cpp:96-147): A1 orbitals reversed, then B1, B2, A2"""
    fx = fixtures.load(os.path.join(GOLDEN, "h2o_631g.npz"))
    L, _, N, _, _ = [int(x) for x in fx["problem/hdr"]]
    ham_irr = [int(i) for i in fx["problem/orb_irrep"]]
    f2 = [h for h in reversed(range(L)) if ham_irr[h] == 0] + [h for ir in (2, 3, 1) for h in range(L) if ham_irr[h] == ir]
    T = fx["problem/tmat"].reshape(L, L, order="F")
    V = fx["problem/vmat"].reshape(L, L, L, L, order="F")
    ix = np.ix_(f2, f2, f2, f2)
    Tp, Vp = T[np.ix_(f2, f2)], V[ix]
    eye = np.eye(L)
    want = Vp + (np.einsum("ac,bd->abcd", eye, Tp) + np.einsum("bd,ac->abcd", eye, Tp)) / (N - 1)
    info, irreps, _, mx = run_host(tmp_path, fx, "c2v")
    assert info["reorder"] == 1
    assert [int(i) for i in irreps] == [ham_irr[h] for h in f2]
    assert np.abs(mx.reshape(L, L, L, L, order="F") - want).max() < 1e-14


@pytest.mark.parametrize("name, reorder, twoS", [("n2_sto3g_singlet", "d2h", 0), ("h2o_631g", "none", 0), ("hubbard10_sextet", "none", 5)])
def test_twodm_and_correlation_accessors(tmp_path, name, reorder, twoS):
    """TwoDM / Correlations accessor arithmetic (TwoDM.cpp:101-231, Correlations.cpp:105-196) restated in numpy on the same filled-in arrays:
    irrep selection rules, 1-RDM and spin-density contractions, trace, energy, entropies, Hamiltonian <-> DMRG orbital order"""
    fx = fixtures.load(os.path.join(GOLDEN, name + ".npz"))
    dump = str(tmp_path / "in.fcidump")
    L, group, N, _, irrep = fcidump_from_fixture(fx, dump)
    out = str(tmp_path / "acc.bin")
    subprocess.run([CALLER, "accessors", dump, str(group), str(twoS), str(N), str(irrep), reorder, out], check=True, capture_output=True, text=True)
    got = np.fromfile(out, dtype=np.float64)
    ham_irr = [int(i) for i in fx["problem/orb_irrep"]]
    if reorder == "d2h":
        f2 = [h for ir in (0, 5, 7, 2, 6, 3, 1, 4) for h in range(L) if ham_irr[h] == ir]
    else:
        f2 = list(range(L))
    f1 = [f2.index(h) for h in range(L)]                      # Hamiltonian -> DMRG
    irr = [ham_irr[f2[d]] for d in range(L)]                  # irreps in DMRG order
    idx = np.arange(L ** 4, dtype=np.float64)
    A = np.sin(0.37 * idx + 0.1).reshape(L, L, L, L, order="F")
    B = np.cos(0.23 * idx - 0.4).reshape(L, L, L, L, order="F")
    ok = np.zeros((L, L, L, L), dtype=bool)
    for a in range(L):
        for b in range(L):
            for c in range(L):
                for d in range(L):
                    ok[a, b, c, d] = (irr[a] ^ irr[b]) == (irr[c] ^ irr[d])
    A, B = np.where(ok, A, 0.0), np.where(ok, B, 0.0)
    tab = [np.sin(1.0 + t + 0.61 * np.arange(L * L)).reshape(L, L, order="F") for t in range(5)]
    T = fx["problem/tmat"].reshape(L, L, order="F")[np.ix_(f2, f2)]
    V = fx["problem/vmat"].reshape(L, L, L, L, order="F")[np.ix_(f2, f2, f2, f2)]
    eye = np.eye(L)
    mx = V + (np.einsum("ac,bd->abcd", eye, T) + np.einsum("bd,ac->abcd", eye, T)) / (N - 1)
    same = np.array([[irr[i] == irr[j] for j in range(L)] for i in range(L)])
    rdm1 = np.where(same, np.einsum("ikjk->ij", A) / (N - 1.0), 0.0)
    spin = np.zeros((L, L))
    if twoS > 0:
        spin = np.where(same, 1.5 * ((2 - N) * rdm1 - np.einsum("ikkj->ij", A) - np.einsum("ikkj->ij", B)) / (0.5 * twoS + 1), 0.0)

    def entropy(i):
        v4 = 0.5 * A[i, i, i, i]; v23 = 0.5 * (rdm1[i, i] - A[i, i, i, i]); v1 = 1.0 - v4 - 2 * v23
        return -sum(m * v * np.log(v) for m, v in ((1, v1), (2, v23), (1, v4)) if v > 1e-100)

    want = [np.einsum("abab->", A), 0.5 * np.sum(A * mx) + float(fx["problem/econst"][0]),
            sum(tab[4][r, c] * abs(r - c) ** 2.0 for r in range(L) for c in range(L) if r != c)]
    for i in range(L):
        want += [entropy(f1[i]), entropy(i)]
    for i in range(L):
        for j in range(L):
            a, b = f1[i], f1[j]
            want += [rdm1[a, b], rdm1[i, j], spin[a, b], spin[i, j], tab[0][a, b], tab[1][a, b], tab[2][a, b], tab[3][a, b], tab[4][a, b], tab[4][i, j],
                     A[a, b, f1[(i + 1) % L], f1[(j + 2) % L]], B[a, b, b, a]]
    want = np.array(want)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-11 * max(1.0, np.abs(want).max())


def test_no_device_aborts_loudly(tmp_path):
    """no CPU fallback: without a CUDA device the DMRG constructor aborts with the library's message"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    fx = fixtures.load(os.path.join(GOLDEN, "n2_sto3g_singlet.npz"))
    dump = str(tmp_path / "in.fcidump")
    L, group, N, twoS, irrep = fcidump_from_fixture(fx, dump)
    res = subprocess.run([CALLER, "solve", dump, str(group), str(twoS), str(N), str(irrep), "none", "50", "0"], capture_output=True, text=True)
    assert res.returncode != 0
    assert "chemps2_b200:" in res.stderr and "failed" in res.stderr


@pytest.mark.gpu
def test_cpp_caller_n2_sto3g_known_answers(tmp_path):
    """The reference's test5 through the C++ mirror: N2/STO-3G 1Ag ground state and first excited state (level shift 20 Eh), D=1000,
    SetupReorderD2h; known answers of tests/test5.cpp.in:83-84 to 1e-8.  The 2-RDM of each state must reproduce its energy and trace."""
    fx = fixtures.load(os.path.join(GOLDEN, "n2_sto3g_singlet.npz"))
    dump = str(tmp_path / "n2.fcidump")
    L, group, N, twoS, irrep = fcidump_from_fixture(fx, dump)
    res = subprocess.run([CALLER, "solve", dump, str(group), str(twoS), str(N), str(irrep), "d2h", "1000", "1"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2JSON ")][-1]
    out = json.loads(line[len("B2JSON "):])
    assert abs(out["energies"][0] - (-107.648250974014)) < 1e-8, out
    assert abs(out["energies"][1] - (-106.944757308768)) < 1e-8, out
    for e, e2 in zip(out["energies"], out["rdm_energies"]):
        assert abs(e - e2) < 1e-7, out
    for tr in out["traces"]:
        assert abs(tr - N * (N - 1)) < 1e-8, out
    assert abs(out["n_elec"] - N) < 1e-8 and out["rdm1_asym"] < 1e-9
    assert out["entropy_sum"] > 0.0 and out["mutinfo_sum"] > 0.0
    assert "Information on completed instruction" in res.stdout


@pytest.mark.parametrize("std", ["c++11", "c++17"])
def test_header_compiles_cleanly(tmp_path, std):
    """the mirror is header-only C++11, warning-free, and its namespace can be renamed for programs that also link the reference"""
    inc = os.path.join(ROOT, "include")
    res = subprocess.run(["g++", f"-std={std}", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", f"-I{inc}", os.path.join(ROOT, "tests", "cpp", "dmrg_caller.cpp")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    src = tmp_path / "ns.cpp"
    src.write_text('#include "chemps2_b200.hpp"\nint main(){ b2shim::ConvergenceScheme s(1); s.setInstruction(0, 10, 1e-8, 3, 0.0);\n'
                   ' b2shim::Irreps g(7); return (s.get_D(0) == 10 && g.getNumberOfIrreps() == 8 && b2shim::Irreps::directProd(5, 7) == 2) ? 0 : 1; }\n')
    exe = tmp_path / "ns"
    lib_dir = os.path.join(ROOT, "chemps2_b200")
    res = subprocess.run(["g++", f"-std={std}", "-DCHEMPS2_B200_NAMESPACE=b2shim", f"-I{inc}", str(src), "-o", str(exe), f"-L{lib_dir}", "-lchemps2_b200",
                          f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    assert subprocess.run([str(exe)]).returncode == 0
