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
