"""CPU tests of the BayesW pieces: the restated ARMS (the source the device compiles) against the reference's own
ARMS object code, and the Gauss-Hermite tables against the exact rule."""
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restated_arms_is_bit_identical_to_the_reference_object_code(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "libarms_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/libarms_ref.so not built (needs /root/reference; oracle/build_ref.sh)")
    exe = str(tmp_path / "arms_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fno-fast-math", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "arms_check.cpp"), "-ldl"], check=True)
    r = subprocess.run([exe, ref, "3000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3000 cases, 0 mismatches" in r.stdout


def test_gauss_hermite_tables_are_the_exact_rule_up_to_the_reference_literals():
    txt = open(os.path.join(ROOT, "hydra_b200", "csrc", "gh_tables.inc")).read()
    assert txt == open(os.path.join(ROOT, "oracle", "gh_tables.h")).read()
    import re
    ns = [int(v) for v in re.search(r"HB_GH_NAME\(n\)\[HB_GH_NRULES\] = \{([^}]*)\}", txt).group(1).split(",")]
    offs = [int(v) for v in re.search(r"HB_GH_NAME\(off\)\[HB_GH_NRULES\] = \{([^}]*)\}", txt).group(1).split(",")]
    xs = [float(v) for v in re.search(r"HB_GH_NAME\(x\)\[HB_GH_NPTS\] = \{([^}]*)\}", txt, re.S).group(1).split(",")]
    ws = [float(v) for v in re.search(r"HB_GH_NAME\(w\)\[HB_GH_NPTS\] = \{([^}]*)\}", txt, re.S).group(1).split(",")]
    assert ns == [3, 5, 7, 9, 11, 13, 15, 17, 25]
    for n, o in zip(ns, offs):
        ex, ew = np.polynomial.hermite.hermgauss(n)
        adj = dict(zip(np.round(ex, 6), ew * np.exp(ex * ex)))
        x, w = np.array(xs[o:o + n - 1]), np.array(ws[o:o + n - 1])
        pos = x[0::2]
        assert np.allclose(sorted(pos), sorted(ex[ex > 1e-9]), rtol=1e-9)
        assert np.allclose(w[0::2], [adj[round(v, 6)] for v in pos], rtol=1e-9)
        asym = [i for i in range(0, n - 1, 2) if x[i + 1] != -x[i]]
        assert asym == ([4] if n >= 11 else []), (n, asym)   # the reference's x6 = -x3 slip, kept for parity


def test_oracle_beta_density_is_log_concave_and_marginals_finite():
    pars = [10.0, 0.01, 3.0, 5000.0, 4000.0, 900.0, 100.0, 0.3, 0.45, 0.3 / 0.45]
    post = oracle.bw_marginal_likelihoods(25, pars, [0.9, 0.05, 0.03, 0.02], [0.001, 0.01, 0.1])
    assert np.all(np.isfinite(post)) and np.all(post > 0)
    if oracle.arms_ref() is not None:
        r = oracle.bw_sample_beta(pars, 0.01, 0.02, 0.0, 1, 0, 0, 0)
        assert r["err"] == 0 and abs(r["beta"]) < 2 * np.sqrt(0.02 * 0.01) and r["nrand"] >= 2
