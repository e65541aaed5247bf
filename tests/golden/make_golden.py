#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference) through oracle/ref_driver.cpp.  Runs only where /root/reference exists
(the build container); the committed .npz files are what travels.

Each sigma case holds: the problem (integral table), bookkeeper dims, every renormalized operator at both boundaries
of a site pair, vec_in / vec_out / diag straight from Heff::makeHeff / Heff::fillHeffDiag (Heff.cpp:43-315), plus the
operator sets before/after DMRG::updateMovingLeft/Right (DMRGoperators.cpp:243-907) for the update parity tests.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from chemps2_b200.fixtures import read_b2fx  # noqa: E402

ME = "/root/reference/tests/matrixelements"
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

# name -> ref_driver arguments.  Systems are the reference's own test systems:
# test1 (N2/STO-3G, D2h, several spin/irrep sectors), test2 (H2O/6-31G, C2v), test3 (CH4/STO-3G), test4 (Hubbard sextet).
CASES = {
    "n2_sto3g_singlet": f"--fcidump {ME}/N2.STO3G.FCIDUMP --group 7 --twoS 0 --N 14 --irrep 0 --D 24 --presweeps 1",
    "n2_sto3g_quintet_b1u": f"--fcidump {ME}/N2.STO3G.FCIDUMP --group 7 --twoS 4 --N 14 --irrep 5 --D 32 --presweeps 0 --reorder",
    "h2o_631g": f"--fcidump {ME}/H2O.631G.FCIDUMP --group 5 --twoS 0 --N 10 --irrep 0 --D 20 --presweeps 1",
    "hubbard10_sextet": "--hubbard 10 4.0 --twoS 5 --N 9 --irrep 0 --D 16 --presweeps 1",
    "ch4_sto3g_triplet_edges": f"--fcidump {ME}/CH4.STO3G.FCIDUMP --group 5 --twoS 2 --N 10 --irrep 1 --D 24 --presweeps 0 --siteA 7 --siteB 0",
    "ch4_sto3g_near_edges": f"--fcidump {ME}/CH4.STO3G.FCIDUMP --group 5 --twoS 0 --N 10 --irrep 0 --D 24 --presweeps 0 --siteA 6 --siteB 1",
    # odd electron number with point-group symmetry: the N2+ cation, doublet Ag, reordered orbitals (half-integer spins in every sector)
    "n2_sto3g_cation_doublet": f"--fcidump {ME}/N2.STO3G.FCIDUMP --group 7 --twoS 1 --N 13 --irrep 0 --D 28 --presweeps 1 --reorder",
    # the reduced BCS (pairing) model of the reference's tests/test12.cpp.in: the folded integral table is written directly with
    # Problem::setMxElement and is NOT 8-fold symmetric (pair scattering <ii|jj> = g without the exchange partners <ij|ji>)
    "pairing8": "--pairing 8 -1.0 0.0 --twoS 0 --N 8 --irrep 0 --D 24 --presweeps 1",
    # the 3 x 3 Hubbard model with periodic boundaries in MOMENTUM space of the reference's tests/test9.cpp.in (doublet, 9 electrons): a dense
    # table with only 4-fold permutation symmetry, again written with Problem::setMxElement
    # (D = 64 after two sweeps: at smaller D the sweep energies of this far-from-converged state move by several 1e-10 when the Davidson tolerance
    # is varied, too close to the 1e-9 parity bar; here the same variation moves them by 2.5e-11)
    "hubbard3x3_momentum": "--hubbard2d 3 5.0 -1.0 --momentum --twoS 1 --N 9 --irrep 0 --D 64 --presweeps 2",
}


# step-by-step traces of noisy sweeps from the seeded random MPS (ref_driver `trace`): the input and every output of Sobject::Split per site
TRACES = {
    "trace_h2o_631g": f"--fcidump {ME}/H2O.631G.FCIDUMP --group 5 --twoS 0 --N 10 --irrep 0 --D 20 --seed 4321 --noise 1e-4 --rtol 1e-10",
    "trace_n2_sto3g_quintet_b1u": f"--fcidump {ME}/N2.STO3G.FCIDUMP --group 7 --twoS 4 --N 14 --irrep 5 --D 14 --seed 77 --noise 3e-4 --rtol 1e-10 --reorder",
    "trace_hubbard10_sextet": "--hubbard 10 4.0 --twoS 5 --N 9 --irrep 0 --D 12 --seed 2024 --noise 1e-4 --rtol 1e-10",
}


def main():
    env = dict(os.environ, OMP_NUM_THREADS="4", OPENBLAS_NUM_THREADS="1")
    only = sys.argv[1:]          # optional: regenerate only the named cases
    for name, args in CASES.items():
        if only and name not in only:
            continue
        tmp = f"/tmp/{name}.b2fx"
        subprocess.run([DRV, "dump", *args.split(), "--seed", "1234", "--out", tmp], check=True, env=env, stdout=subprocess.DEVNULL)
        fx = read_b2fx(tmp)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")
    for name, args in TRACES.items():
        if only and name not in only:
            continue
        tmp = f"/tmp/{name}.b2fx"
        subprocess.run([DRV, "trace", *args.split(), "--out", tmp], check=True, env=env, stdout=subprocess.DEVNULL)
        fx = read_b2fx(tmp)
        for k in [k for k in fx if k in ("problem/vmat", "problem/tmat")]:
            del fx[k]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")
    if only:
        return
    # problem-only fixture of BASELINE config 2 (N2/cc-pVDZ, 14e/28o, D2h, SetupReorderD2h): the input of scripts/run_dmrg.py n2_ccpvdz,
    # whose known answers are the published energies of sphinx/resources.rst:50-56.  Only the folded table gMxElement is kept.
    tmp = "/tmp/n2_ccpvdz_problem.b2fx"
    subprocess.run([DRV, "problem", "--fcidump", f"{ME}/N2.CCPVDZ.FCIDUMP", "--group", "7", "--twoS", "0", "--N", "14", "--irrep", "0", "--reorder",
                    "--out", tmp], check=True, env=env, stdout=subprocess.DEVNULL)
    fx = read_b2fx(tmp)
    np.savez_compressed(os.path.join(HERE, "problem_n2_ccpvdz.npz"), **{k: v for k, v in fx.items() if k not in ("problem/vmat", "problem/tmat")})
    print("problem_n2_ccpvdz", os.path.getsize(os.path.join(HERE, "problem_n2_ccpvdz.npz")) // 1024, "KiB")
    tmp = "/tmp/wigner.b2fx"
    subprocess.run([DRV, "wigner", "--out", tmp], check=True, env=env)
    np.savez_compressed(os.path.join(HERE, "wigner.npz"), **read_b2fx(tmp))


if __name__ == "__main__":
    main()
