"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference (oracle/_ref/ref_driver, built by oracle/build_ref.sh from /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may import this module; the product
package chemps2_b200/ never does.
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def available():
    return os.path.exists(REF_DRIVER)


def run_reference_synth(w, seed, reps=1, amp=1.0, threads=None, out_path=None, workdir="/tmp", dims=None):
    """Runs the UNMODIFIED reference's Heff::makeHeff (oracle/_ref/ref_driver synth) on workload `w` with hash-filled
    operators.  Test / bench-baseline infrastructure only.  -> dict(mean_s, best_s, threads, veclength, vec_out, diag)"""
    if not os.path.exists(REF_DRIVER):
        raise FileNotFoundError(REF_DRIVER + " missing (oracle/build_ref.sh builds it where /root/reference exists)")
    pfile = os.path.join(workdir, f"b2_problem_{w.name}_{os.getpid()}.bin")
    ofile = out_path or os.path.join(workdir, f"b2_refout_{w.name}_{os.getpid()}.bin")
    w.write_problem_file(pfile)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count())
    cmd = [REF_DRIVER, "synth", "--problem", pfile, "--D", str(w.D), "--site", str(w.site), "--reps", str(reps), "--seed", str(seed),
           "--amp", repr(amp), "--out", ofile]
    dfile = None
    if dims is not None:
        dfile = os.path.join(workdir, f"b2_dims_{w.name}_{os.getpid()}.bin")
        np.ascontiguousarray(dims, dtype="<i4").tofile(dfile)
        cmd += ["--dims", dfile]
    res = subprocess.run(cmd, env=env, check=True, capture_output=True, text=True)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2REF synth")][-1].split()
    kv = {line[i]: line[i + 1] for i in range(2, len(line) - 1, 2)}
    raw = np.fromfile(ofile, dtype="<f8")
    n = int(kv["veclength"])
    os.remove(pfile)
    if dfile:
        os.remove(dfile)
    if out_path is None:
        os.remove(ofile)
    return dict(mean_s=float(kv["mean_s"]), best_s=float(kv["best_s"]), diag_s=float(kv["diag_s"]), setup_s=float(kv["setup_s"]),
                threads=int(kv["threads"]), veclength=n, vec_out=raw[:n], diag=raw[n:2 * n])


def op_key(side, kind, i, j):
    """the key of oracle/ref_driver.cpp op_key / chemps2_b200's b2_opset_fill_hash"""
    return (side << 60) | (kind << 40) | ((i + 1) << 20) | (j + 1)


def run_reference_update(w, seed, moving_right=True, site=None, dims=None, amp=1.0, amp_t=0.1, threads=None, workdir="/tmp"):
    """Runs the UNMODIFIED reference's DMRG::updateMovingRight / updateMovingLeft (ref_driver `synthupdate`) on hash-filled operators of
    workload `w`: -> dict(update_s, threads, ops = [(kind, i, j, size, sum, sumsq, dot)]) with dot = <operator, hash(seed + 17, key(5, kind, i, j))>"""
    from chemps2_b200.fixtures import read_b2fx
    if not os.path.exists(REF_DRIVER):
        raise FileNotFoundError(REF_DRIVER)
    tag = f"{w.name}_{os.getpid()}"
    pfile, ofile, dfile = (os.path.join(workdir, f"b2_{k}_{tag}.bin") for k in ("uproblem", "uout", "udims"))
    w.write_problem_file(pfile)
    np.ascontiguousarray(dims, dtype="<i4").tofile(dfile)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads or os.cpu_count()))
    cmd = [REF_DRIVER, "synthupdate", "--problem", pfile, "--site", str(w.site if site is None else site), "--moving-right", "1" if moving_right else "0",
           "--seed", str(seed), "--amp", repr(amp), "--amp-t", repr(amp_t), "--dims", dfile, "--out", ofile]
    res = subprocess.run(cmd, env=env, check=True, capture_output=True, text=True)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("B2REF synthupdate")][-1].split()
    kv = {line[i]: line[i + 1] for i in range(2, len(line) - 1, 2)}
    fx = read_b2fx(ofile)
    meta, sums = fx["upd/meta"].reshape(-1, 4), fx["upd/sums"].reshape(-1, 3)
    for f in (pfile, ofile, dfile):
        os.remove(f)
    return dict(update_s=float(kv["update_s"]), setup_s=float(kv["setup_s"]), threads=int(kv["threads"]),
                ops=[(int(m[0]), int(m[1]), int(m[2]), int(m[3]), float(s[0]), float(s[1]), float(s[2])) for m, s in zip(meta, sums)])
