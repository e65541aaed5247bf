"""CheMPS2::Davidson as a reverse-communication object with device vectors (b2_davidson_*; Davidson.h:46-58): driven exactly like
Heff::SolveDAVIDSON_main drives the reference's class (Heff.cpp:331-386) — 'A' guess + diagonal, 'B' matrix-vector products,
'C' result — it must give the eigenvalue of the fused b2_heff_solve and of the reference's own sweep record."""
import ctypes as C

import numpy as np
import pytest

import cpu_check
from chemps2_b200._lib import check, lib, vp

pytestmark = pytest.mark.gpu


class _Dev:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}


@pytest.mark.parametrize("tag", ["A", "B"])
def test_reverse_communication_davidson(golden, tag):
    import torch
    ctx, left, right, heff = cpu_check.build_case(golden, tag, device=0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    site = int(golden[tag + "/hdr"][0])
    labels, offs = ctx.sobject_table(site)
    scale = np.concatenate([np.full(offs[k + 1] - offs[k], np.sqrt(labels[k][7] + 1.0)) for k in range(len(labels))])
    guess = golden[tag + "/joined"] * scale                      # prog2symm (Sobject.cpp:624-636)
    n = heff.n
    d = vp()
    check(lib.b2_davidson_create(ctx.h, n, 32, 3, 1e-8, 1e-12, C.byref(d)))   # Options.h:70-72 + the sweep's rtol
    try:
        instr, p0, p1 = C.c_char(), vp(), vp()
        check(lib.b2_davidson_fetch(d, C.byref(instr), C.byref(p0), C.byref(p1)))
        assert instr.value == b"A"
        torch.as_tensor(_Dev(p0.value, n), device="cuda").copy_(torch.from_numpy(guess))
        torch.as_tensor(_Dev(p1.value, n), device="cuda").copy_(torch.from_numpy(heff.diag()))
        torch.cuda.synchronize()
        nmult = 0
        while True:
            check(lib.b2_davidson_fetch(d, C.byref(instr), C.byref(p0), C.byref(p1)))
            if instr.value != b"B":
                break
            heff.apply_device(p0.value, p1.value)
            nmult += 1
            assert nmult < 500
        assert instr.value == b"C"
        e_rc = lib.b2_davidson_eigenvalue(d)
        assert lib.b2_davidson_num_multiplications(d) == nmult > 0
        x = torch.as_tensor(_Dev(p0.value, n), device="cuda").cpu().numpy().copy()
        assert abs(float(torch.as_tensor(_Dev(p1.value, 1), device="cuda").cpu()[0]) - e_rc) == 0.0
    finally:
        lib.b2_davidson_destroy(d)
    e_fused, sol, nm = heff.solve(golden[tag + "/joined"], rtol=1e-8)
    assert abs(e_rc - e_fused) < 1e-10 and nmult == nm
    assert abs(np.linalg.norm(x) - 1.0) < 1e-10
    assert np.abs(np.abs(x) - np.abs(sol * scale)).max() < 1e-7
    # the reference's own record of this micro-iteration (DMRG::solve_site energies of the golden sweep)
    L = ctx.L
    en = golden["energies"]
    npre = len(en) - 2 * (L - 2)
    idx = npre + ((L - 2 - site) if tag == "A" else (L - 2) + site)
    assert abs(e_rc + float(golden["problem/econst"][0]) - en[idx]) < 1e-9
