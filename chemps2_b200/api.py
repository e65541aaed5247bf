"""Thin object wrappers over the C ABI; names mirror the reference classes they stand for
(Context ~ Problem + SyBookkeeper, OpSet ~ the operator tables of DMRG.h:211-226, Heff ~ CheMPS2::Heff)."""
import ctypes as C

import numpy as np

from ._lib import B2Error, FlatPresum, FlatTerm, c_dp, c_ip, c_lp, check, lib, vp  # noqa: F401


def _dp(a):
    return a.ctypes.data_as(c_dp)


class Context:
    def __init__(self, device=0):
        self.h = vp()
        check(lib.b2_ctx_create(int(device), C.byref(self.h)))
        self.L = 0

    def close(self):
        if self.h:
            lib.b2_ctx_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_problem(self, L, group, N, twoS, irrep, orb_irrep, mx=None, tmat=None, vmat=None, econst=0.0):
        irr = np.ascontiguousarray(orb_irrep, dtype=np.int32)
        self.L = int(L)
        if mx is not None:
            mx = np.ascontiguousarray(mx, dtype=np.float64)
            check(lib.b2_problem_set(self.h, L, group, N, twoS, irrep, irr.ctypes.data_as(c_ip), _dp(mx), float(econst)))
        else:
            t = np.ascontiguousarray(tmat, dtype=np.float64)
            v = np.ascontiguousarray(vmat, dtype=np.float64)
            check(lib.b2_problem_set_integrals(self.h, L, group, N, twoS, irrep, irr.ctypes.data_as(c_ip), _dp(t), _dp(v), float(econst)))

    def set_option(self, name, value):
        check(lib.b2_ctx_set_option(self.h, name.encode(), float(value)))

    def set_stream(self, cuda_stream):
        check(lib.b2_ctx_set_stream(self.h, vp(int(cuda_stream))))

    def mx_elem(self):
        out = np.zeros(self.L ** 4, dtype=np.float64)
        check(lib.b2_problem_mx(self.h, _dp(out)))
        return out

    def bk_init(self, D):
        check(lib.b2_bk_init(self.h, int(D)))

    def bk_import(self, rows):
        """rows: int array (n, 6) of boundary, N, twoS, irrep, curdim, fcidim as dumped by the reference"""
        for b, n, ts, ir, cur, _ in np.asarray(rows).reshape(-1, 6):
            check(lib.b2_bk_set_dim(self.h, int(b), int(n), int(ts), int(ir), int(cur)))

    def dim(self, b, n, ts, ir):
        return lib.b2_bk_dim(self.h, b, n, ts, ir)

    def fcidim(self, b, n, ts, ir):
        return lib.b2_bk_fcidim(self.h, b, n, ts, ir)

    def sobject_table(self, site):
        nk = lib.b2_sobject_nkappa(self.h, site)
        labels = np.zeros((nk, 9), dtype=np.int32)
        offs = np.zeros(nk + 1, dtype=np.int64)
        check(lib.b2_sobject_table(self.h, site, labels.ctypes.data_as(c_ip), offs.ctypes.data_as(c_lp)))
        return labels, offs


def twodm_fill_site(ctx, site, t_storage, left, right, A=None, B=None):
    """TwoDM::FillSite on the GPU: -> (A, B) flat L^4 arrays (created zeroed when not given); see b2_twodm_fill_site"""
    n = ctx.L ** 4
    A = np.zeros(n) if A is None else A
    B = np.zeros(n) if B is None else B
    t = np.ascontiguousarray(t_storage, dtype=np.float64)
    check(lib.b2_twodm_fill_site(ctx.h, int(site), _dp(t), left.h if left else None, right.h if right else None, _dp(A), _dp(B)))
    return A, B


def svd_batch(ctx, mats):
    """thin SVDs of a list of 2-D arrays on the GPU -> [(u, s, vt)]   (b2_svd_batch; stands for dgesdd_ in Sobject::Split)"""
    As = [np.asfortranarray(m, dtype=np.float64) for m in mats]
    ms = np.array([a.shape[0] for a in As], dtype=np.int32)
    ns = np.array([a.shape[1] for a in As], dtype=np.int32)
    ks = np.minimum(ms, ns)
    S = [np.zeros(k) for k in ks]
    U = [np.zeros((m, k), order="F") for m, k in zip(ms, ks)]
    VT = [np.zeros((k, n), order="F") for k, n in zip(ks, ns)]
    arr = lambda xs: (c_dp * max(len(xs), 1))(*[x.ctypes.data_as(c_dp) for x in xs])   # noqa: E731
    check(lib.b2_svd_batch(ctx.h, len(As), ms.ctypes.data_as(c_ip), ns.ctypes.data_as(c_ip), arr(As), arr(S), arr(U), arr(VT)))
    return list(zip(U, S, VT))


class OpSet:
    def __init__(self, ctx, boundary, moving_right, correlation=False):
        """correlation=True: the {G, Y, Z, K, M} helper tensors of DMRG::update_correlations_tensors instead of the sweep operators"""
        self.ctx = ctx
        self.h = vp()
        if correlation:
            check(lib.b2_opset_create_correlation(ctx.h, int(boundary), C.byref(self.h)))
        else:
            check(lib.b2_opset_create(ctx.h, int(boundary), int(bool(moving_right)), C.byref(self.h)))

    def close(self):
        if self.h:
            lib.b2_opset_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib.b2_opset_count(self.h)

    def info(self, idx):
        k, i, j, s = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        check(lib.b2_opset_info(self.h, idx, C.byref(k), C.byref(i), C.byref(j), C.byref(s)))
        return k.value, i.value, j.value, s.value

    def find(self, kind, i, j):
        return lib.b2_opset_find(self.h, kind, i, j)

    def upload(self, idx, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        assert a.size == self.info(idx)[3], (a.size, self.info(idx))
        check(lib.b2_opset_upload(self.h, idx, _dp(a)))

    def download(self, idx):
        a = np.zeros(self.info(idx)[3], dtype=np.float64)
        check(lib.b2_opset_download(self.h, idx, _dp(a)))
        return a

    def upload_all(self, ops, strict=True):
        """ops: [(kind, i, j, data)] as returned by fixtures.split_ops"""
        for kind, i, j, data in ops:
            idx = self.find(kind, i, j)
            if idx < 0:
                if strict:
                    raise KeyError(f"operator kind={kind} ({i},{j}) not in this set")
                continue
            self.upload(idx, data)

    def offload(self):
        check(lib.b2_opset_offload(self.h))

    def offload_file(self, path):
        """park the arena in a file (NVMe tier); reload() brings it back and removes the file"""
        check(lib.b2_opset_offload_file(self.h, str(path).encode()))

    def reload(self):
        check(lib.b2_opset_reload(self.h))

    def resident(self):
        return bool(lib.b2_opset_resident(self.h))

    def fill_hash(self, seed, amp=1.0):
        check(lib.b2_opset_fill_hash(self.h, int(seed), float(amp)))

    def host_arena(self):
        n = lib.b2_opset_arena_size(self.h)
        p = lib.b2_opset_host_arena(self.h)
        return np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0)


class Heff:
    """sigma = H_eff * S for the site pair (site, site+1); stands for CheMPS2::Heff (Heff.h:50-70)."""

    def __init__(self, ctx, site, left, right, world=1, rank=0):
        self.ctx, self.left, self.right = ctx, left, right
        self.h = vp()
        check(lib.b2_heff_create(ctx.h, int(site), left.h if left else None, right.h if right else None, int(world), int(rank), C.byref(self.h)))
        self.n = lib.b2_heff_veclength(self.h)
        self._site = int(site)

    def close(self):
        if self.h:
            lib.b2_heff_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, vec):
        v = np.ascontiguousarray(vec, dtype=np.float64)
        out = np.empty_like(v)
        check(lib.b2_heff_apply(self.h, _dp(v), _dp(out)))
        return out

    def apply_device(self, dev_in_ptr, dev_out_ptr):
        check(lib.b2_heff_apply_device(self.h, vp(dev_in_ptr), vp(dev_out_ptr)))

    def set_excitations(self, vectors):
        """vectors: list of VeffTilde arrays (symmetric convention); [] switches the projector off (Heff.h:70 nLower / VeffTilde)"""
        vs = [np.ascontiguousarray(v, dtype=np.float64) for v in vectors]
        arr = (c_dp * max(len(vs), 1))(*[_dp(v) for v in vs])
        check(lib.b2_heff_set_excitations(self.h, len(vs), arr))

    def kernel_seconds(self):
        return lib.b2_heff_last_kernel_seconds(self.h)

    def diag(self):
        out = np.zeros(self.n, dtype=np.float64)
        check(lib.b2_heff_diag(self.h, _dp(out)))
        return out

    def solve(self, s_prog, rtol=1e-5):
        """Heff::SolveDAVIDSON: s_prog = Sobject storage (program convention) -> (eigenvalue, solution, n_matvec)"""
        v = np.array(s_prog, dtype=np.float64, copy=True)
        e, nm = C.c_double(), C.c_int()
        check(lib.b2_heff_solve(self.h, _dp(v), float(rtol), C.byref(e), C.byref(nm)))
        return e.value, v, nm.value

    def stats(self):
        o = np.zeros(12)
        check(lib.b2_heff_stats(self.h, _dp(o)))
        keys = ["terms", "terms_zero", "presums", "flops_ref", "flops_exec", "work_doubles", "stage1", "tiles", "waves", "launches",
                "part_doubles", "worklist_bytes"]
        return {k: float(v) for k, v in zip(keys, o)}

    def export(self):
        nt = lib.b2_heff_num_terms(self.h)
        terms = (FlatTerm * max(nt, 1))()
        check(lib.b2_heff_export_terms(self.h, terms))
        npp = lib.b2_heff_num_presum_parts(self.h)
        parts = (FlatPresum * max(npp, 1))()
        check(lib.b2_heff_export_presums(self.h, parts))
        return terms, nt, parts, npp, lib.b2_heff_presum_size(self.h)


class Update:
    """operator renormalisation of one sweep step; stands for DMRG::updateMovingRight/Left (DMRGoperators.cpp:243-907)"""

    def __init__(self, ctx, index, moving_right, old_set, new_set, world=1, rank=0):
        self.ctx, self.old_set, self.new_set = ctx, old_set, new_set
        self.h = vp()
        check(lib.b2_update_create_sharded(ctx.h, int(index), int(bool(moving_right)), old_set.h if old_set else None, new_set.h,
                                           int(world), int(rank), C.byref(self.h)))

    def set_allreduce(self, allreduce):
        """allreduce: an AllReduce object (kept alive by this Update)"""
        self._allreduce = allreduce
        check(lib.b2_update_set_allreduce(self.h, allreduce.cfn, None))

    def close(self):
        if self.h:
            lib.b2_update_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, t_storage):
        t = np.ascontiguousarray(t_storage, dtype=np.float64)
        check(lib.b2_update_run(self.h, _dp(t)))

    def run_device(self, t_dev_ptr):
        """T already resident on the device (asynchronous on the context stream)"""
        check(lib.b2_update_run_device(self.h, vp(t_dev_ptr)))

    def stats(self):
        o = np.zeros(8)
        check(lib.b2_update_stats(self.h, _dp(o)))
        keys = ["terms", "mix_terms", "presums", "flops_ref", "flops_exec", "work_doubles", "waves", "launches"]
        return {k: float(v) for k, v in zip(keys, o)}


class DMRG:
    """sweep driver; stands for the hot-path part of CheMPS2::DMRG (DMRG.cpp:357-452)"""

    def __init__(self, ctx):
        self.ctx = ctx
        self.h = vp()
        check(lib.b2_dmrg_create(ctx.h, C.byref(self.h)))

    def close(self):
        if self.h:
            lib.b2_dmrg_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_world(self, world, rank, allreduce=None):
        """shard the sweep over `world` GPUs; allreduce: an AllReduce object (kept alive by this driver)"""
        self._allreduce = allreduce
        check(lib.b2_dmrg_set_world(self.h, int(world), int(rank), allreduce.cfn if allreduce else None, None))

    def save_mps(self, path, converged=False):
        check(lib.b2_dmrg_save_mps(self.h, str(path).encode(), int(bool(converged))))

    def load_mps(self, path):
        """-> converged flag stored in the checkpoint; operator sets are dropped (call presolve())"""
        c = C.c_int()
        check(lib.b2_dmrg_load_mps(self.h, str(path).encode(), C.byref(c)))
        return bool(c.value)

    def set_plan_cache(self, enabled):
        check(lib.b2_dmrg_set_plan_cache(self.h, int(bool(enabled))))

    def plan_cache_stats(self):
        hits, misses = C.c_longlong(), C.c_longlong()
        check(lib.b2_dmrg_plan_cache_stats(self.h, C.byref(hits), C.byref(misses)))
        return hits.value, misses.value

    def set_plan_prefetch(self, enabled):
        check(lib.b2_dmrg_set_plan_prefetch(self.h, int(bool(enabled))))

    def plan_prefetched(self):
        return int(lib.b2_dmrg_plan_prefetched(self.h))

    def presolve(self):
        check(lib.b2_dmrg_presolve(self.h))

    def solve(self, scheme):
        """DMRG::Solve; scheme = [(D, energy_conv, max_sweeps, noise_prefactor, davidson_rtol), ...] like ConvergenceScheme::set_instruction"""
        Ds = np.array([x[0] for x in scheme], dtype=np.int32)
        ec = np.array([x[1] for x in scheme], dtype=np.float64)
        ms = np.array([x[2] for x in scheme], dtype=np.int32)
        nz = np.array([x[3] for x in scheme], dtype=np.float64)
        rt = np.array([x[4] for x in scheme], dtype=np.float64)
        e = C.c_double()
        check(lib.b2_dmrg_solve(self.h, len(scheme), Ds.ctypes.data_as(c_ip), _dp(ec), ms.ctypes.data_as(c_ip), _dp(nz), _dp(rt), C.byref(e)))
        return e.value

    def new_excitation(self, eshift, D, seed):
        """DMRG::newExcitation: store the current MPS as a lower state (level shift eshift), restart from a random MPS at dimension D"""
        check(lib.b2_dmrg_new_excitation(self.h, float(eshift), int(D), int(seed)))

    def calc_2rdm(self):
        """-> (A, B) as [L,L,L,L] arrays indexed [i,j,k,l] (TwoDM::getTwoDMA_DMRG / getTwoDMB_DMRG); see b2_dmrg_calc_2rdm"""
        L = self.ctx.L
        A, B = np.zeros(L ** 4), np.zeros(L ** 4)
        check(lib.b2_dmrg_calc_2rdm(self.h, _dp(A), _dp(B)))
        return A.reshape((L, L, L, L), order="F"), B.reshape((L, L, L, L), order="F")

    def calc_correlations(self, A, B):
        """A, B from calc_2rdm -> dict of [L,L] tables Cspin, Cdens, Cspinflip, Cdirad, MutInfo (Correlations::get*_DMRG)"""
        L = self.ctx.L
        a = np.ascontiguousarray(A.ravel(order="F"))
        b = np.ascontiguousarray(B.ravel(order="F"))
        out = {k: np.zeros(L * L) for k in ("Cspin", "Cdens", "Cspinflip", "Cdirad", "MutInfo")}
        check(lib.b2_dmrg_calc_correlations(self.h, _dp(a), _dp(b), *[_dp(out[k]) for k in ("Cspin", "Cdens", "Cspinflip", "Cdirad", "MutInfo")]))
        return {k: v.reshape((L, L), order="F") for k, v in out.items()}

    def set_spill(self, enabled, directory=None):
        """spill mode: only the operator sets in use stay in HBM; directory: park the others in files there instead of pinned host memory"""
        check(lib.b2_dmrg_set_spill_dir(self.h, str(directory).encode() if directory else None))
        check(lib.b2_dmrg_set_spill(self.h, int(bool(enabled))))

    def timers(self, reset=False):
        o = np.zeros(5)
        check(lib.b2_dmrg_timers(self.h, _dp(o), int(bool(reset))))
        return dict(plan_s=o[0], solve_s=o[1], split_s=o[2], update_s=o[3], n_matvec=int(o[4]))

    def set_mps(self, site, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        check(lib.b2_dmrg_set_mps(self.h, int(site), _dp(a)))

    def get_mps(self, site):
        a = np.zeros(lib.b2_dmrg_mps_size(self.h, int(site)), dtype=np.float64)
        check(lib.b2_dmrg_get_mps(self.h, int(site), _dp(a)))
        return a

    def random_mps(self, seed):
        check(lib.b2_dmrg_random_mps(self.h, int(seed)))

    def srand(self, seed):
        """re-seed the rand() stream that feeds the noise (a run started from a checkpoint)"""
        check(lib.b2_dmrg_srand(self.h, int(seed)))

    def update(self, index, moving_right):
        check(lib.b2_dmrg_update(self.h, int(index), int(bool(moving_right))))

    def opset_download(self, boundary, moving_right, kind, i, j):
        """one operator of the driver's set at (boundary, direction) -> numpy, or None"""
        h = lib.b2_dmrg_opset(self.h, int(boundary), int(bool(moving_right)))
        if not h:
            return None
        idx = lib.b2_opset_find(h, kind, i, j)
        if idx < 0:
            return None
        size = C.c_int64()
        check(lib.b2_opset_info(h, idx, None, None, None, C.byref(size)))
        a = np.zeros(size.value, dtype=np.float64)
        check(lib.b2_opset_download(h, idx, _dp(a)))
        return a

    def solve_site(self, index, rtol, noise, D, moving_right, change):
        e, dw, nm = C.c_double(), C.c_double(), C.c_int()
        check(lib.b2_dmrg_solve_site(self.h, int(index), float(rtol), float(noise), int(D), int(bool(moving_right)), int(bool(change)),
                                     C.byref(e), C.byref(dw), C.byref(nm)))
        return e.value, dw.value, nm.value

    def sweep_info(self):
        out = (C.c_double * 4)()
        check(lib.b2_dmrg_sweep_info(self.h, out))
        return dict(last_energy=out[0], last_min_energy=out[1], max_discarded=out[2], total_min_energy=out[3])

    def sweep(self, to_right, rtol, noise, D, change):
        e, dw = C.c_double(), C.c_double()
        check(lib.b2_dmrg_sweep(self.h, int(bool(to_right)), float(rtol), float(noise), int(D), int(bool(change)), C.byref(e), C.byref(dw)))
        return e.value, dw.value


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, vp, vp, C.c_int64, vp)


SVD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(c_dp), C.POINTER(c_dp), C.POINTER(c_dp),
                     C.POINTER(c_dp))


def _lapack_svd(user, count, m, n, a, sv, u, vt):
    """a caller-side SVD for b2_sobject_split (numpy -> LAPACK gesdd), the convention of b2_svd_batch: thin factors, column-major"""
    try:
        for i in range(count):
            mi, ni = m[i], n[i]
            k = min(mi, ni)
            A = np.ctypeslib.as_array(a[i], shape=(mi * ni,)).reshape(mi, ni, order="F")
            U, S, VT = np.linalg.svd(A, full_matrices=False)
            np.ctypeslib.as_array(sv[i], shape=(k,))[:] = S
            np.ctypeslib.as_array(u[i], shape=(mi * k,))[:] = U.ravel(order="F")
            np.ctypeslib.as_array(vt[i], shape=(k * ni,))[:] = VT.ravel(order="F")
        return 0
    except Exception:
        return 1


LAPACK_SVD = SVD_FN(_lapack_svd)


def split(ctx, site, s_storage, D, moving_right, change, svd=None):
    """Sobject::Split through b2_sobject_split -> (t_left, t_right, discarded weight); svd = None: batched device SVD, or an SVD_FN
    (LAPACK_SVD keeps the decomposition on the host).  The bookkeeper of ctx holds the new dimensions of boundary site+1 afterwards."""
    sv = np.ascontiguousarray(s_storage, dtype=np.float64)
    r, dw = C.c_void_p(), C.c_double()
    fn = C.cast(svd, C.c_void_p) if svd is not None else None
    check(lib.b2_sobject_split(ctx.h, int(site), _dp(sv), int(D), int(bool(moving_right)), int(bool(change)), fn, None, C.byref(r), C.byref(dw)))
    try:
        out = []
        for which in (0, 1):
            t = np.zeros(max(int(lib.b2_split_size(r, which)), 1), dtype=np.float64)
            check(lib.b2_split_get(r, which, _dp(t)))
            out.append(t[:int(lib.b2_split_size(r, which))])
    finally:
        lib.b2_split_destroy(r)
    return out[0], out[1], dw.value


class Join:
    """Sobject::Join (Sobject.cpp:212-258) of the site tensors of (site, site+1) for the current bookkeeper dimensions"""

    def __init__(self, ctx, site):
        self.ctx, self.site = ctx, int(site)
        self.h = C.c_void_p()
        check(lib.b2_join_create(ctx.h, self.site, C.byref(self.h)))
        self.n = int(lib.b2_sobject_size(ctx.h, self.site))

    def close(self):
        if self.h:
            lib.b2_join_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, t_left, t_right):
        tl = np.ascontiguousarray(t_left, dtype=np.float64)
        tr = np.ascontiguousarray(t_right, dtype=np.float64)
        out = np.zeros(max(self.n, 1), dtype=np.float64)
        check(lib.b2_join_run(self.h, _dp(tl), _dp(tr), _dp(out)))
        return out[:self.n]


class AllReduce:
    """b2_allreduce_fn backed by torch.distributed (NCCL on the GPU box): sums a device vector over the ranks in place on the
    stream the library hands over (the context stream = torch's current stream in bench.py).  Stands in for MPI_Allreduce of
    the reference (Heff.cpp:350-365, DMRGoperators.cpp:449-533)."""

    class _Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.calls, self.doubles = 0, 0

        def fn(user, ptr, n, stream):
            try:
                t = torch.as_tensor(AllReduce._Dev(ptr, n), device="cuda")
                # run the collective ON the stream the library hands over: the kernels that produced the vector and the
                # ones that consume it are ordered on that stream, not on torch's current one
                if stream:
                    with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                        dist.all_reduce(t)
                else:
                    dist.all_reduce(t)
                self.calls += 1
                self.doubles += int(n)
                return 0
            except Exception as e:   # never let an exception cross the C boundary
                print("AllReduce callback failed:", e)
                return -1

        self.cfn = ALLREDUCE_FN(fn)


def context_from_fixture(fx, tag, device=-1):
    """Build a Context (problem + bookkeeper dims) from a golden fixture section `tag` ('A' or 'B')."""
    L, group, N, twoS, irrep = [int(x) for x in fx["problem/hdr"]]
    ctx = Context(device)
    ctx.set_problem(L, group, N, twoS, irrep, fx["problem/orb_irrep"], mx=fx["problem/mx"], econst=float(fx["problem/econst"][0]))
    ctx.bk_init(1)
    ctx.bk_import(fx[tag + "/bk"])
    return ctx


KIND_NAMES = ["L", "S0", "S1", "F0", "F1", "A", "B", "C", "D", "Q", "X", "G", "Y", "Z", "K", "M"]


S_KEY = (3 << 60)   # key of the synthetic two-site vector, same as op_key(3, 0, -1, -1) in oracle/ref_driver.cpp


def hash_fill(n, seed, key=S_KEY, amp=1.0):
    out = np.zeros(int(n), dtype=np.float64)
    check(lib.b2_hash_fill(_dp(out), int(n), int(seed), int(key), float(amp)))
    return out
