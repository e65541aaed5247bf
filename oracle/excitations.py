"""TEST INFRASTRUCTURE (oracle): numpy restatement of the excited-state level-shift projector of the reference.

Heff::addDiagramExcitations (HeffDiagrams1.cpp:65-85): for every lower state s, alpha = ddot(S, VeffTilde[s]) over the whole
two-site vector and sigma[block] += alpha * VeffTilde[s][block]  ==>  sigma += sum_s <V_s|S> V_s.
Heff::addDiagonalExcitations (HeffDiagonal.cpp:621-640): diag[i] += VeffTilde[s][i]^2.

Pinned by the golden fixtures (tests/golden/*.npz keys A|B/exc_*: Heff::makeHeff / fillHeffDiag of the unmodified reference called
with nLower = 2).  Only tests/ may import this module; the product path is b2_heff_set_excitations (CUDA)."""
import numpy as np


def add_excitations(sigma_plain, vec_in, veff_tilde):
    out = np.array(sigma_plain, dtype=np.float64, copy=True)
    for v in veff_tilde:
        out += float(np.dot(vec_in, v)) * v
    return out


def add_diagonal_excitations(diag_plain, veff_tilde):
    out = np.array(diag_plain, dtype=np.float64, copy=True)
    for v in veff_tilde:
        out += v * v
    return out
