"""The C++ host: hydra's command line and file formats on top of the C ABI.
CPU part: option parsing / input readers (--dry-run). GPU part: the reference's own test strategy:
run-vs-run equality -- BED input vs sparse input vs the Python front-end, compared through the output files."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import random_bed, simulate_y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "hydra_b200", "bin", "hydra_b200")


def _exe():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", os.path.join(ROOT, "hydra_b200", "csrc")], check=True, capture_output=True)
        subprocess.run(["make", "-C", os.path.join(ROOT, "hydra_b200", "host")], check=True, capture_output=True)
    return EXE


def write_dataset(d, N=600, M=150, G=2, n_na=7, seed=3):
    rng = np.random.default_rng(seed)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    with open(os.path.join(d, "t.bed"), "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(bed.tobytes())
    with open(os.path.join(d, "t.bim"), "w") as f:
        for j in range(M):
            f.write(f"1\trs{j}\t0\t{j + 1}\tA\tC\n")
    with open(os.path.join(d, "t.fam"), "w") as f:
        for i in range(N):
            f.write(f"F{i} I{i} 0 0 1 -9\n")
    y = simulate_y(rng, g)
    na = np.sort(rng.choice(N, n_na, replace=False))
    with open(os.path.join(d, "t.phen"), "w") as f:
        for i in range(N):
            f.write(f"F{i} I{i} {'NA' if i in set(na.tolist()) else repr(float(y[i]))}\n")
    groups = (np.arange(M) % G).astype(np.int32)
    with open(os.path.join(d, "t.group"), "w") as f:
        for j in range(M):
            f.write(f"rs{j} {groups[j]}\n")
    with open(os.path.join(d, "t.mS"), "w") as f:
        f.write(";".join(["0.001,0.01,0.1"] * G) + "\n")
    return bed, y, na.astype(np.uint32), groups


def read_bet(path, M, dtype=np.float64):
    raw = open(path, "rb").read()
    (m,) = struct.unpack("<I", raw[:4])
    assert m == M
    rec = 4 + M * np.dtype(dtype).itemsize
    n = (len(raw) - 4) // rec
    its, vals = [], []
    for r in range(n):
        o = 4 + r * rec
        its.append(struct.unpack("<I", raw[o:o + 4])[0])
        vals.append(np.frombuffer(raw[o + 4:o + rec], dtype=dtype))
    return np.array(its), np.array(vals)


def base_args(d, out, extra=()):
    return [_exe(), "--mpibayes", "bayesMPI", "--pheno", os.path.join(d, "t.phen"), "--number-individuals", "600", "--number-markers", "150",
            "--chain-length", "6", "--thin", "2", "--save", "4", "--seed", "1222", "--shuf-mark", "1", "--sync-rate", "5", "--tasks", "3",
            "--groupIndexFile", os.path.join(d, "t.group"), "--groupMixtureFile", os.path.join(d, "t.mS"),
            "--mcmc-out-dir", os.path.join(d, out), "--mcmc-out-name", "run", *extra]


def test_cli_rejects_unknown_and_incomplete_options(tmp_path):
    r = subprocess.run([_exe(), "--no-such-flag"], capture_output=True, text=True)
    assert r.returncode != 0 and "invalid option" in r.stderr
    r = subprocess.run([_exe(), "--mpibayes", "bayesMPI", "--mcmc-out-dir", str(tmp_path), "--mcmc-out-name", "x"], capture_output=True, text=True)
    assert r.returncode != 0 and "--number-individuals" in r.stderr
    r = subprocess.run([_exe(), "--mpibayes", "bayesMPI", "--sparse-dir", "x", "--mcmc-out-dir", str(tmp_path), "--mcmc-out-name", "x",
                        "--number-individuals", "5", "--number-markers", "5"], capture_output=True, text=True)
    assert r.returncode != 0 and "--sparse-dir and --sparse-basename" in r.stderr


def test_cli_dry_run_reads_reference_formats(tmp_path):
    d = str(tmp_path)
    write_dataset(d)
    r = subprocess.run(base_args(d, "o", ["--bfile", os.path.join(d, "t"), "--dry-run"]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "N = 600 (7 NA phenotypes), M = 150, 3 task(s), sync rate 5, 2 group(s) x 3 mixtures, input bed" in r.stdout
    # a wrong --number-markers is caught against the .bim file
    a = base_args(d, "o", ["--bfile", os.path.join(d, "t"), "--dry-run"])
    a[a.index("--number-markers") + 1] = "151"
    r = subprocess.run(a, capture_output=True, text=True)
    assert r.returncode != 0 and ".bim file has 150 markers" in r.stdout


def test_cli_accepts_the_reference_sync_options_and_refuses_unsupported_ones(tmp_path):
    d = str(tmp_path)
    write_dataset(d)
    r = subprocess.run(base_args(d, "o", ["--bfile", os.path.join(d, "t"), "--dry-run", "--sparse-sync", "--bed-sync", "--ignore-xfiles"]),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("has no effect here") == 3
    r = subprocess.run(base_args(d, "o", ["--bfile", os.path.join(d, "t"), "--dry-run", "--check-RAM"]), capture_output=True, text=True)
    assert r.returncode != 0 and "not supported by hydra_b200" in r.stdout + r.stderr
    # --covariates: phenotype and covariate files are read together, an NA in either drops the individual (src/data.cpp:1615-1672)
    rng = np.random.default_rng(0)
    with open(os.path.join(d, "t.cov"), "w") as f:
        for i in range(600):
            f.write(f"F{i} I{i} {rng.normal():.6f} {'NA' if i == 11 else f'{rng.normal():.6f}'}\n")
    r = subprocess.run(base_args(d, "o", ["--bfile", os.path.join(d, "t"), "--dry-run", "--covariates", os.path.join(d, "t.cov")]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "numFixedEffect = 2" in r.stdout and "(8 NA phenotypes)" in r.stdout    # 7 NA phenotypes + 1 NA covariate


def test_cli_dry_run_bayesfh_options_and_prior_files(tmp_path):
    """--mpibayes bayesFHMPI with its five options (src/options.cpp:116-135) and the two prior files in the reference's ';' / ','
    text format (src/data.cpp:2034-2096) are parsed and validated without a GPU."""
    d = str(tmp_path)
    write_dataset(d)
    open(os.path.join(d, "t.gp"), "w").write("4.0,0.2; 2.5,0.05\n")
    open(os.path.join(d, "t.dp"), "w").write("5,1,1,1;\n 1,2,2,0.5\n")

    def run(extra, typ="bayesFHMPI"):
        a = base_args(d, "o", ["--bfile", os.path.join(d, "t"), "--dry-run"] + extra)
        a[a.index("bayesMPI")] = typ
        return subprocess.run(a, capture_output=True, text=True)
    r = run(["--groupPriorsFile", os.path.join(d, "t.gp"), "--dPriorsFile", os.path.join(d, "t.dp"), "--tau0", "0.5", "--v0t", "4", "--v0L", "2.5"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bayesFH hyper-parameters: tau0 0.5 v0t 4 v0c 3 s02c 1 v0L 2.5" in r.stdout
    assert "group priors (v0G, s02G) of 2 group(s)" in r.stdout and "Dirichlet parameters of 2 group(s) x 4 components" in r.stdout
    r = run(["--tau0", "-1"])
    assert r.returncode != 0 and "must be positive" in r.stdout + r.stderr
    open(os.path.join(d, "bad.dp"), "w").write("5,1,1; 1,2,2\n")          # one value short per group
    r = run(["--dPriorsFile", os.path.join(d, "bad.dp")], typ="bayesMPI")
    assert r.returncode != 0 and "3 values, 4 needed" in r.stdout + r.stderr
    r = run(["--groupPriorsFile", os.path.join(d, "nofile.gp")], typ="bayesMPI")
    assert r.returncode != 0 and "can not open the --groupPriorsFile file" in r.stdout + r.stderr
    r = run([], typ="bayesXYZ")
    assert r.returncode != 0 and "bayesFHMPI (BayesFH)" in r.stdout + r.stderr


def test_cli_reads_the_reference_example_files():
    # parser fixtures shipped with the reference (SURVEY 8c iv): copied line counts only, no reference file is read at GPU time
    ex = "/root/reference/example"
    if not os.path.isdir(ex):
        pytest.skip("reference checkout not present")
    r = subprocess.run([_exe(), "--mpibayes", "bayesMPI", "--sparse-dir", "/tmp", "--sparse-basename", "none", "--pheno", ex + "/normal.phen",
                        "--number-individuals", "5000", "--number-markers", "10000", "--groupIndexFile", ex + "/normal.group",
                        "--groupMixtureFile", ex + "/normal.mS", "--mcmc-out-dir", "/tmp/hb_o", "--mcmc-out-name", "x", "--dry-run"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "N = 5000 (0 NA phenotypes), M = 10000" in r.stdout and "2 group(s) x 3 mixtures" in r.stdout


@pytest.mark.gpu
def test_cli_bed_vs_sparse_vs_python_runs_are_identical(tmp_path):
    import hydra_b200
    d = str(tmp_path)
    bed, y, na, groups = write_dataset(d)
    N, M = 600, 150
    r = subprocess.run(base_args(d, "bed", ["--bfile", os.path.join(d, "t")]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # bed -> sparse files -> sparse-input run
    r = subprocess.run([_exe(), "--bed-to-sparse", "--bfile", os.path.join(d, "t"), "--sparse-dir", os.path.join(d, "sp"), "--sparse-basename", "t",
                        "--number-individuals", str(N), "--number-markers", str(M)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(os.path.join(d, "sp", "t.dim")).read().split() == [str(N), str(M)]
    import oracle
    sp = oracle.sparse_fill_indices(bed, N)   # the sparse files hold the raw (not NA-corrected) lists, ascending, absolute starts
    for ext, dt, want in (("si1", np.uint32, sp.I1), ("ss1", np.uint64, sp.N1S), ("sl1", np.uint64, sp.N1L),
                          ("si2", np.uint32, sp.I2), ("ss2", np.uint64, sp.N2S), ("sl2", np.uint64, sp.N2L),
                          ("sim", np.uint32, sp.IM), ("ssm", np.uint64, sp.NMS), ("slm", np.uint64, sp.NML)):   # all nine files of data.cpp:1196-1221
        assert np.array_equal(np.fromfile(os.path.join(d, "sp", "t." + ext), dt), want), ext
    r = subprocess.run(base_args(d, "sparse", ["--sparse-dir", os.path.join(d, "sp"), "--sparse-basename", "t"]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for ext in ("bet", "cpn", "acu", "csv", "xbet", "xcpn", "eps.0", "eps.2", "mrk.1", "mus.0"):
        a = open(os.path.join(d, "bed", "run." + ext), "rb").read()
        b = open(os.path.join(d, "sparse", "run." + ext), "rb").read()
        # the kernel's sums are exact integers (fixed-point epsilon): list and 2-bit records give the same bytes everywhere
        assert a == b, ext
    its, beta = read_bet(os.path.join(d, "bed", "run.bet"), M)
    its2, beta2 = read_bet(os.path.join(d, "sparse", "run.bet"), M)
    assert its.tolist() == [0, 2, 4] and its2.tolist() == [0, 2, 4]
    assert np.array_equal(beta, beta2)
    # the same chain through the Python front-end
    keep = np.setdiff1d(np.arange(N), na)
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=3, sync_rate=5, n_groups=2, n_mix=4, repr_mode="bed") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y[keep], [[0.001, 0.01, 0.1]] * 2, groups=groups, seed=1222)
        for it in range(5):
            brr.iteration()
            if it % 2 == 0:
                b, c, _ = brr.state()
                assert np.array_equal(b, beta[it // 2]), f"python vs CLI beta at iteration {it}"
    _, comp = read_bet(os.path.join(d, "bed", "run.cpn"), M, np.int32)
    assert comp.shape == (3, M) and comp.min() >= 0 and comp.max() <= 3
    lines = open(os.path.join(d, "bed", "run.csv")).read().split("\n")[:-1]
    assert len(lines) == 3 and len({len(l) for l in lines}) == 1   # fixed line length (the reference seeks by n*strlen)
    assert len(lines[0].split(",")) == 2 + 2 + 5 + 8
    raw = open(os.path.join(d, "bed", "run.eps.0"), "rb").read()
    assert struct.unpack("<II", raw[:8]) == (4, N - len(na)) and len(raw) == 8 + 8 * (N - len(na))
    _check_with_reference_converters(os.path.join(d, "bed", "run"), M, N - len(na), its, beta, comp)


def _check_with_reference_converters(base, M, Nc, its, beta, comp):
    """The reference's OWN readers of its output files (postproc/beta_converter.cpp:33-52, components_converter.cpp:33-52,
    epsilon_converter.cpp:30-43), compiled from where they lie by oracle/build_ref.sh, must read the files this host writes:
    record offsets, iteration headers and every value (format oracles). Skipped only where oracle/_ref was
    never built (no /root/reference)."""
    import re
    conv = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(conv, "beta_converter")):
        pytest.skip("oracle/_ref converters not built")
    n_rec = len(its)
    out = subprocess.run([os.path.join(conv, "beta_converter"), base + ".bet", str(n_rec - 1)], capture_output=True, text=True).stdout
    assert f"{M} markers were processed." in out
    heads = [int(x) for x in re.findall(r"read iteration number (\d+) \(iter=", out)]
    assert heads == list(its)
    vals = re.findall(r"^\s*(\d+)/\s*(\d+) =\s*(\S+)$", out, re.M)
    assert len(vals) == n_rec * M
    got = np.array([float(v[2]) for v in vals]).reshape(n_rec, M)
    assert [int(v[0]) for v in vals[::M]] == list(its) and [int(v[1]) for v in vals[:M]] == list(range(M))
    np.testing.assert_allclose(got, beta, rtol=0, atol=0.6e-12)     # the converter prints %20.12f
    out = subprocess.run([os.path.join(conv, "components_converter"), base + ".cpn", str(n_rec - 1)], capture_output=True, text=True).stdout
    assert f"{M} markers were processed." in out
    # (the reference's components converter passes a double to %2d, so only its record headers are meaningful)
    assert [int(x) for x in re.findall(r"read iteration number (\d+) \(iter=", out)] == list(its)
    assert len(re.findall(r"^\s*\d+/\s*\d+ =", out, re.M)) == n_rec * M and comp.shape == (n_rec, M)
    out = subprocess.run([os.path.join(conv, "epsilon_converter"), base + ".eps.0"], capture_output=True, text=True).stdout
    assert "iteration 4 was last logged into epsilon file." in out and f"{Nc} individuals were processed." in out
    ev = np.array([float(x) for x in re.findall(r"^\s*4/\s*\d+ =\s*(\S+)$", out, re.M)])
    eps = np.frombuffer(open(base + ".eps.0", "rb").read()[8:], np.float64)
    assert len(ev) == Nc
    np.testing.assert_allclose(ev, eps, rtol=0, atol=0.6e-11)       # %20.11f


@pytest.mark.gpu
def test_cli_restart_continues_the_chain_bit_for_bit(tmp_path):
    """hydra's --restart (src/BayesRRm.cpp:842-928): a run interrupted after a --save point and continued with --restart
    writes the same files, byte for byte, as the uninterrupted run (the chain state file keeps exact doubles and the
    random streams, where the reference re-reads 15 digits of its .csv)."""
    d = str(tmp_path)
    write_dataset(d)
    chain = ["--chain-length", "9", "--thin", "1", "--save", "4"]
    r = subprocess.run(base_args(d, "full", ["--bfile", os.path.join(d, "t")] + chain), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    cut = ["--chain-length", "6", "--thin", "1", "--save", "4"]   # stops after iteration 5, last save point = iteration 4
    r = subprocess.run(base_args(d, "part", ["--bfile", os.path.join(d, "t")] + cut), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run(base_args(d, "part", ["--bfile", os.path.join(d, "t"), "--restart"] + chain), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "restarting after iteration 4" in r.stdout
    for ext in ("csv", "bet", "cpn", "acu", "xbet", "xcpn", "eps.0", "eps.2", "mrk.1", "mus.0", "mus.2", "rst.0"):
        a = open(os.path.join(d, "full", "run." + ext), "rb").read()
        b = open(os.path.join(d, "part", "run." + ext), "rb").read()
        assert a == b, ext
    # a state file of another problem is refused
    r = subprocess.run(base_args(d, "part", ["--bfile", os.path.join(d, "t"), "--restart", "--tasks", "2"] + chain), capture_output=True, text=True)
    assert r.returncode != 0 and "FATAL" in r.stdout


def test_cli_reads_the_weibull_example_files():
    ex = "/root/reference/example"
    if not os.path.isdir(ex):
        pytest.skip("reference checkout not present")
    r = subprocess.run([_exe(), "--mpibayes", "bayesWMPI", "--sparse-dir", "/tmp", "--sparse-basename", "none", "--pheno", ex + "/Weibull.phen",
                        "--failure", ex + "/Weibull.fail", "--quad_points", "25", "--number-individuals", "5000", "--number-markers", "10000",
                        "--S", "0.001,0.01,0.1", "--mcmc-out-dir", "/tmp/hb_o", "--mcmc-out-name", "x", "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bayesWMPI, N = 5000 (0 NA phenotypes), M = 10000" in r.stdout


@pytest.mark.gpu
def test_cli_bayesw_run_equals_python_run(tmp_path):
    import hydra_b200
    d = str(tmp_path)
    N, M = 600, 80
    rng = np.random.default_rng(21)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    with open(os.path.join(d, "w.bed"), "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + bed.tobytes())
    open(os.path.join(d, "w.bim"), "w").write("".join(f"1\trs{j}\t0\t{j + 1}\tA\tC\n" for j in range(M)))
    open(os.path.join(d, "w.fam"), "w").write("".join(f"F{i} I{i} 0 0 1 -9\n" for i in range(N)))
    y = 4.1 + 0.1 * np.log(rng.exponential(size=N))
    fail = (rng.random(N) > 0.1).astype(int)
    na_p, na_f = {5, 17}, {40}
    open(os.path.join(d, "w.phen"), "w").write("".join(f"F{i} I{i} {'NA' if i in na_p else repr(float(y[i]))}\n" for i in range(N)))
    open(os.path.join(d, "w.fail"), "w").write("".join(f"{-9 if i in na_f else fail[i]}\n" for i in range(N)))
    r = subprocess.run([_exe(), "--mpibayes", "bayesWMPI", "--bfile", os.path.join(d, "w"), "--pheno", os.path.join(d, "w.phen"),
                        "--failure", os.path.join(d, "w.fail"), "--quad_points", "11", "--number-individuals", str(N), "--number-markers", str(M),
                        "--S", "0.001,0.01,0.1", "--chain-length", "4", "--thin", "1", "--save", "2", "--seed", "9", "--sync-rate", "3", "--tasks", "2",
                        "--mcmc-out-dir", os.path.join(d, "o"), "--mcmc-out-name", "w"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    its, beta = read_bet(os.path.join(d, "o", "w.bet"), M)
    assert its.tolist() == [0, 1, 2, 3]
    na = np.array(sorted(na_p | na_f), np.uint32)
    keep = np.setdiff1d(np.arange(N), na)
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=2, sync_rate=3, n_groups=1, n_mix=4, repr_mode="bed", model="bayesW") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        bw = hydra_b200.BayesW(st, y[keep], fail[keep].astype(float), [[0.001, 0.01, 0.1]], quad_points=11, seed=9)
        for it in range(4):
            o = bw.iteration()
            assert np.array_equal(bw.state()[0], beta[it])
    line = open(os.path.join(d, "o", "w.csv")).read().split("\n")[3].split(",")
    assert int(line[0]) == 3 and abs(float(line[1]) - o["mu"]) < 1e-12 and abs(float(line[3]) - o["alpha"]) < 1e-12
    # --restart: stop after iteration 2 (its --save point), continue to 4 iterations: the same files byte for byte
    base = [_exe(), "--mpibayes", "bayesWMPI", "--bfile", os.path.join(d, "w"), "--pheno", os.path.join(d, "w.phen"),
            "--failure", os.path.join(d, "w.fail"), "--quad_points", "11", "--number-individuals", str(N), "--number-markers", str(M),
            "--S", "0.001,0.01,0.1", "--thin", "1", "--save", "2", "--seed", "9", "--sync-rate", "3", "--tasks", "2",
            "--mcmc-out-dir", os.path.join(d, "p"), "--mcmc-out-name", "w"]
    r = subprocess.run(base + ["--chain-length", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run(base + ["--chain-length", "4", "--restart"], capture_output=True, text=True)
    assert r.returncode == 0 and "restarting after iteration 2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    for ext in ("csv", "bet", "cpn", "eps.0"):
        assert open(os.path.join(d, "o", "w." + ext), "rb").read() == open(os.path.join(d, "p", "w." + ext), "rb").read(), ext


@pytest.mark.gpu
def test_cli_two_processes_two_gpus_equal_one_process(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = str(tmp_path)
    write_dataset(d)
    common = ["--bfile", os.path.join(d, "t")]
    a1 = base_args(d, "one", common)
    a1[a1.index("--tasks") + 1] = "4"
    r = subprocess.run(a1, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    procs = []
    for rank in range(2):
        a2 = base_args(d, "two", common + ["--rank", str(rank), "--world", "2"])
        a2[a2.index("--tasks") + 1] = "4"
        procs.append(subprocess.Popen(a2, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                      env={**os.environ, "MASTER_PORT": "31999", "HB_PEER_TIMEOUT_S": "30"}))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs[0][-1500:] + outs[1][-1500:]
    M = 150
    i1, b1 = read_bet(os.path.join(d, "one", "run.bet"), M)
    i2, b2 = read_bet(os.path.join(d, "two", "run.bet"), M)
    assert i1.tolist() == i2.tolist() == [0, 2, 4]
    np.testing.assert_allclose(b2, b1, rtol=1e-9, atol=1e-14)
    assert open(os.path.join(d, "one", "run.cpn"), "rb").read() == open(os.path.join(d, "two", "run.cpn"), "rb").read()
    for t in range(4):
        assert os.path.exists(os.path.join(d, "two", f"run.mus.{t}")) and os.path.exists(os.path.join(d, "two", f"run.eps.{t}"))
    e1 = np.fromfile(os.path.join(d, "one", "run.eps.3"), np.float64, offset=8)
    e2 = np.fromfile(os.path.join(d, "two", "run.eps.3"), np.float64, offset=8)
    np.testing.assert_allclose(e2, e1, rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_cli_covariates_run_equals_python_run(tmp_path):
    """hydra's --covariates through the C++ host: same chain as the Python front-end; .gam / .xiv written at the save points
    in the reference's layout (u32 it; u32 len; f64 / i32 [len], src/BayesRRm.cpp:2811-2831)."""
    import hydra_b200
    d = str(tmp_path)
    bed, y, na, groups = write_dataset(d)
    N, M, F = 600, 150, 2
    rng = np.random.default_rng(4)
    X = rng.normal(size=(N, F))
    with open(os.path.join(d, "t.cov"), "w") as f:
        for i in range(N):
            f.write(f"F{i} I{i} " + " ".join(repr(float(v)) for v in X[i]) + "\n")
    r = subprocess.run(base_args(d, "cov", ["--bfile", os.path.join(d, "t"), "--covariates", os.path.join(d, "t.cov")]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    its, beta = read_bet(os.path.join(d, "cov", "run.bet"), M)
    keep = np.setdiff1d(np.arange(N), na)
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=3, sync_rate=5, n_groups=2, n_mix=4, repr_mode="bed") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y[keep], [[0.001, 0.01, 0.1]] * 2, groups=groups, seed=1222, covariates=X[keep])
        for it in range(5):
            brr.iteration()
            if it % 2 == 0:
                assert np.array_equal(brr.state()[0], beta[it // 2]), f"python vs CLI beta at iteration {it}"
            if it == 4:
                gam, xiv = brr.gamma()
    raw = open(os.path.join(d, "cov", "run.gam.0"), "rb").read()
    assert struct.unpack("<II", raw[:8]) == (4, F) and np.array_equal(np.frombuffer(raw[8:], np.float64), gam)
    raw = open(os.path.join(d, "cov", "run.xiv.0"), "rb").read()
    assert struct.unpack("<II", raw[:8]) == (4, F) and np.array_equal(np.frombuffer(raw[8:], np.int32), xiv)


@pytest.mark.gpu
def test_cli_bayesw_covariates_run_equals_python_run(tmp_path):
    """--covariates with bayesWMPI (src/BayesW.cpp:1366-1413): phenotype, covariate and failure files read together
    (src/data.cpp:1679-1752: an NA in any of them drops the individual); .gam as text lines "%5d, %20.17f, ..." per thinned
    iteration (:1970-1980), .xiv at the save points; --restart continues byte for byte."""
    import hydra_b200
    d = str(tmp_path)
    N, M, F = 600, 80, 2
    rng = np.random.default_rng(22)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    with open(os.path.join(d, "w.bed"), "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + bed.tobytes())
    open(os.path.join(d, "w.bim"), "w").write("".join(f"1\trs{j}\t0\t{j + 1}\tA\tC\n" for j in range(M)))
    open(os.path.join(d, "w.fam"), "w").write("".join(f"F{i} I{i} 0 0 1 -9\n" for i in range(N)))
    X = np.column_stack([rng.integers(0, 2, N).astype(np.float64), rng.normal(size=N)])
    y = 4.1 + X @ np.array([0.04, -0.02]) + 0.1 * np.log(rng.exponential(size=N))
    fail = (rng.random(N) > 0.1).astype(int)
    na_p, na_f, na_c = {5, 17}, {40}, {77}
    open(os.path.join(d, "w.phen"), "w").write("".join(f"F{i} I{i} {'NA' if i in na_p else repr(float(y[i]))}\n" for i in range(N)))
    open(os.path.join(d, "w.fail"), "w").write("".join(f"{-9 if i in na_f else fail[i]}\n" for i in range(N)))
    open(os.path.join(d, "w.cov"), "w").write("".join(f"F{i} I{i} {'NA' if i in na_c else repr(float(X[i, 0]))} {repr(float(X[i, 1]))}\n" for i in range(N)))
    base = [_exe(), "--mpibayes", "bayesWMPI", "--bfile", os.path.join(d, "w"), "--pheno", os.path.join(d, "w.phen"),
            "--failure", os.path.join(d, "w.fail"), "--covariates", os.path.join(d, "w.cov"), "--quad_points", "11",
            "--number-individuals", str(N), "--number-markers", str(M), "--S", "0.001,0.01,0.1", "--thin", "1", "--save", "2",
            "--seed", "9", "--sync-rate", "3", "--tasks", "2", "--mcmc-out-name", "w"]
    r = subprocess.run(base + ["--chain-length", "4", "--mcmc-out-dir", os.path.join(d, "o")], capture_output=True, text=True)
    assert r.returncode == 0 and "numFixedEffect = 2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    its, beta = read_bet(os.path.join(d, "o", "w.bet"), M)
    na = np.array(sorted(na_p | na_f | na_c), np.uint32)
    keep = np.setdiff1d(np.arange(N), na)
    gam_lines = open(os.path.join(d, "o", "w.gam")).read().strip().split("\n")
    assert len(gam_lines) == 4
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=2, sync_rate=3, n_groups=1, n_mix=4, repr_mode="bed", model="bayesW") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        bw = hydra_b200.BayesW(st, y[keep], fail[keep].astype(float), [[0.001, 0.01, 0.1]], quad_points=11, seed=9, covariates=X[keep])
        for it in range(4):
            bw.iteration()
            assert np.array_equal(bw.state()[0], beta[it])
            gam, xiv = bw.gamma()
            f = gam_lines[it].split(",")
            assert int(f[0]) == it and np.allclose([float(v) for v in f[1:]], gam, rtol=0, atol=1e-16)
            if it == 2:
                raw = open(os.path.join(d, "o", "w.xiv"), "rb").read()
                assert struct.unpack("<II", raw[:8]) == (2, F) and np.array_equal(np.frombuffer(raw[8:], np.int32), xiv)
    assert np.abs(gam).max() > 1e-3
    r = subprocess.run(base + ["--chain-length", "3", "--mcmc-out-dir", os.path.join(d, "p")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run(base + ["--chain-length", "4", "--mcmc-out-dir", os.path.join(d, "p"), "--restart"], capture_output=True, text=True)
    assert r.returncode == 0 and "restarting after iteration 2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    for ext in ("csv", "bet", "cpn", "gam", "eps.0"):
        assert open(os.path.join(d, "o", "w." + ext), "rb").read() == open(os.path.join(d, "p", "w." + ext), "rb").read(), ext


@pytest.mark.gpu
def test_cli_restart_from_the_output_files(tmp_path):
    """The reference's own restart (src/BayesRRm.cpp:842-928): without a .rst state file the chain state is read back from the
    output files of the last save point (.csv, .xbet, .xcpn, .mus, .eps, .mrk -- reference layouts). The random streams are
    re-seeded, so the continuation is checked against the Python front-end restored from the same files."""
    import glob
    import hydra_b200
    d = str(tmp_path)
    bed, y, na, groups = write_dataset(d)
    N, M, T = 600, 150, 3
    def args(n):
        a = base_args(d, "o", ["--bfile", os.path.join(d, "t")])
        a[a.index("--chain-length") + 1] = str(n)
        a[a.index("--thin") + 1] = "1"
        return a
    r = subprocess.run(args(5), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    o = os.path.join(d, "o", "run")
    for f in glob.glob(o + ".rst.*"):
        os.remove(f)
    its0, beta0 = read_bet(o + ".bet", M)
    assert its0.tolist() == [0, 1, 2, 3, 4]
    # what the files of the save point (iteration 4) hold
    line = [float(v) for v in open(o + ".csv").read().strip().split("\n")[4].split(",")]
    G, K = 2, 4
    assert int(line[0]) == 4 and int(line[1]) == G
    sigmaG, sigmaE, pi = np.array(line[2:2 + G]), line[2 + G], np.array(line[2 + G + 5:]).reshape(G, K)
    xb = open(o + ".xbet", "rb").read(); xc = open(o + ".xcpn", "rb").read()
    assert struct.unpack("<II", xb[:8]) == (M, 4)
    beta4, comp4 = np.frombuffer(xb[8:], np.float64), np.frombuffer(xc[8:], np.int32)
    assert np.array_equal(beta4, beta0[4])
    mu = np.array([np.frombuffer(open(f"{o}.mus.{t}", "rb").read(), np.dtype([("it", "<u4"), ("mu", "<f8")]))["mu"][4] for t in range(T)])
    eps0 = np.frombuffer(open(o + ".eps.0", "rb").read()[8:], np.float64)
    perm = np.concatenate([np.frombuffer(open(f"{o}.mrk.{t}", "rb").read()[8:], np.int32) for t in range(T)])
    r = subprocess.run(args(8) + ["--restart"], capture_output=True, text=True)
    assert r.returncode == 0 and "iteration_to_restart_from = 4" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    its, beta = read_bet(o + ".bet", M)
    assert its.tolist() == list(range(8)) and np.array_equal(beta[:5], beta0)          # the old records stay, three are appended
    assert len(open(o + ".csv").read().strip().split("\n")) == 8
    keep = np.setdiff1d(np.arange(N), na)
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=T, sync_rate=5, n_groups=G, n_mix=K, repr_mode="bed") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y[keep], [[0.001, 0.01, 0.1]] * G, groups=groups, seed=1222)
        brr.restore_outputs(5, sigmaG, pi, sigmaE, mu, beta4, comp4, eps0, perm)
        for it in range(5, 8):
            brr.iteration()
            assert np.array_equal(brr.state()[0], beta[it]), f"python (restored from the files) vs CLI --restart at iteration {it}"
    assert not np.array_equal(beta[5], beta[4])


@pytest.mark.gpu
def test_cli_bayesfh_and_prior_files_equal_python_run(tmp_path):
    """--mpibayes bayesFHMPI with --tau0/--v0t/--v0c/--s02c/--v0L (src/options.cpp:116-135) and the two prior files
    (--groupPriorsFile "v0,s02; v0,s02", --dPriorsFile "a,b,c,d; ...", src/data.cpp:2034-2096) through the C++ host: same chain as
    the Python front-end, and --restart (state file) continues it byte for byte."""
    import hydra_b200
    d = str(tmp_path)
    bed, y, na, groups = write_dataset(d)
    N, M = 600, 150
    open(os.path.join(d, "t.gp"), "w").write("4.0,0.2; 2.5,0.05\n")
    open(os.path.join(d, "t.dp"), "w").write("5,1,1,1; 1,2,2,0.5\n")
    fhargs = ["--bfile", os.path.join(d, "t"), "--dPriorsFile", os.path.join(d, "t.dp"), "--tau0", "0.5", "--v0t", "4", "--v0L", "2.5"]

    def args(out, extra):
        a = base_args(d, out, fhargs + extra)
        a[a.index("bayesMPI")] = "bayesFHMPI"
        return a
    r = subprocess.run(args("fh", []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bayesFH: tau0 0.5 v0t 4 v0c 3 s02c 1 v0L 2.5" in r.stdout
    its, beta = read_bet(os.path.join(d, "fh", "run.bet"), M)
    keep = np.setdiff1d(np.arange(N), na)
    with hydra_b200.GenotypeStore(N, M, na_inds=na, tasks=3, sync_rate=5, n_groups=2, n_mix=4, repr_mode="bed") as st:
        st.load_data_from_bed(bed)
        st.finalize()
        brr = hydra_b200.BayesRRm(st, y[keep], [[0.001, 0.01, 0.1]] * 2, groups=groups, seed=1222, fh=dict(tau0=0.5, v0t=4.0, v0L=2.5),
                                  dirichlet_priors=[[5, 1, 1, 1], [1, 2, 2, 0.5]])
        for it in range(5):
            brr.iteration()
            if it % 2 == 0:
                assert np.array_equal(brr.state()[0], beta[it // 2]), f"python vs CLI beta at iteration {it}"
        assert (beta[-1] != 0).any()
    # restart from the state file of the save point (iteration 4): a run of 10 equals a run of 6 continued to 10
    a10 = args("fh10", [])
    a10[a10.index("--chain-length") + 1] = "10"
    assert subprocess.run(a10, capture_output=True, text=True).returncode == 0
    ar = args("fh", ["--restart"])
    ar[ar.index("--chain-length") + 1] = "10"
    r = subprocess.run(ar, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for ext in (".bet", ".cpn", ".csv"):
        assert open(os.path.join(d, "fh", "run" + ext), "rb").read() == open(os.path.join(d, "fh10", "run" + ext), "rb").read(), ext
    # --groupPriorsFile with bayesMPI: the sigmaG column of the .csv differs from the run with the built-in priors, pi does too
    r1 = subprocess.run(base_args(d, "gp", ["--bfile", os.path.join(d, "t"), "--groupPriorsFile", os.path.join(d, "t.gp")]), capture_output=True, text=True)
    r0 = subprocess.run(base_args(d, "gp0", ["--bfile", os.path.join(d, "t")]), capture_output=True, text=True)
    assert r1.returncode == 0 and r0.returncode == 0, r1.stdout[-1500:] + r1.stderr[-1500:]
    c1, c0 = open(os.path.join(d, "gp", "run.csv")).read(), open(os.path.join(d, "gp0", "run.csv")).read()
    assert c1 != c0 and c1.split("\n")[0].split(",")[0] == c0.split("\n")[0].split(",")[0]
    # malformed prior file: loud
    open(os.path.join(d, "bad.gp"), "w").write("4.0,0.2\n")
    r = subprocess.run(base_args(d, "bad", ["--bfile", os.path.join(d, "t"), "--groupPriorsFile", os.path.join(d, "bad.gp")]), capture_output=True, text=True)
    assert r.returncode != 0 and "1 groups, the run has 2" in r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_bayesw_two_processes_two_gpus_equal_one_process(tmp_path):
    """bayesWMPI with one process per GPU (mpirun -np 2 hydra --mpibayes bayesWMPI): the two processes' .bet / .cpn slices, rank 0's
    .csv and residual equal the one-process run with the same 4 tasks (1e-9: the sum of the epsilon changes is taken in another
    order); --restart from the two .rst.<rank> files continues byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = str(tmp_path)
    N, M = 600, 80
    rng = np.random.default_rng(21)
    bed, g = random_bed(rng, M, N, pmiss=0.01)
    with open(os.path.join(d, "w.bed"), "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + bed.tobytes())
    open(os.path.join(d, "w.bim"), "w").write("".join(f"1\trs{j}\t0\t{j + 1}\tA\tC\n" for j in range(M)))
    open(os.path.join(d, "w.fam"), "w").write("".join(f"F{i} I{i} 0 0 1 -9\n" for i in range(N)))
    x = np.where(g < 0, 0, g).astype(np.float64)
    y = 4.1 + 0.02 * (x[3] - x[3].mean()) - 0.03 * (x[50] - x[50].mean()) + 0.1 * np.log(rng.exponential(size=N))
    fail = (rng.random(N) > 0.1).astype(int)
    open(os.path.join(d, "w.phen"), "w").write("".join(f"F{i} I{i} {'NA' if i in (5, 17) else repr(float(y[i]))}\n" for i in range(N)))
    open(os.path.join(d, "w.fail"), "w").write("".join(f"{fail[i]}\n" for i in range(N)))

    def args(out, extra):
        return [_exe(), "--mpibayes", "bayesWMPI", "--bfile", os.path.join(d, "w"), "--pheno", os.path.join(d, "w.phen"),
                "--failure", os.path.join(d, "w.fail"), "--quad_points", "11", "--number-individuals", str(N), "--number-markers", str(M),
                "--S", "0.001,0.01,0.1", "--thin", "1", "--save", "2", "--seed", "9", "--sync-rate", "3", "--tasks", "4",
                "--mcmc-out-dir", os.path.join(d, out), "--mcmc-out-name", "w", *extra]

    def run2(out, extra, port):
        procs = [subprocess.Popen(args(out, extra + ["--rank", str(r), "--world", "2"]), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                  env={**os.environ, "MASTER_PORT": port}) for r in range(2)]
        outs = [p.communicate(timeout=300)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), outs[0][-1500:] + outs[1][-1500:]
        return outs
    r = subprocess.run(args("one", ["--chain-length", "5"]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    run2("two", ["--chain-length", "5"], "31997")
    i1, b1 = read_bet(os.path.join(d, "one", "w.bet"), M)
    i2, b2 = read_bet(os.path.join(d, "two", "w.bet"), M)
    assert i1.tolist() == i2.tolist() == [0, 1, 2, 3, 4] and (b1 != 0).any()
    np.testing.assert_allclose(b2, b1, rtol=1e-8, atol=1e-13)
    assert open(os.path.join(d, "one", "w.cpn"), "rb").read() == open(os.path.join(d, "two", "w.cpn"), "rb").read()
    e1 = np.fromfile(os.path.join(d, "one", "w.eps.0"), np.float64, offset=8)
    e2 = np.fromfile(os.path.join(d, "two", "w.eps.0"), np.float64, offset=8)
    np.testing.assert_allclose(e2, e1, rtol=1e-9, atol=1e-11)
    c1 = [float(v) for v in open(os.path.join(d, "one", "w.csv")).read().split("\n")[4].split(",")]
    c2 = [float(v) for v in open(os.path.join(d, "two", "w.csv")).read().split("\n")[4].split(",")]
    np.testing.assert_allclose(c2, c1, rtol=1e-9)
    # restart on two GPUs: 3 iterations (save point at 2), then continued to 5 = the uninterrupted two-GPU run, byte for byte
    run2("twor", ["--chain-length", "3"], "31996")
    outs = run2("twor", ["--chain-length", "5", "--restart"], "31995")
    assert "restarting after iteration 2" in outs[0]
    for ext in ("csv", "bet", "cpn", "eps.0"):
        assert open(os.path.join(d, "two", "w." + ext), "rb").read() == open(os.path.join(d, "twor", "w." + ext), "rb").read(), ext


@pytest.mark.gpu
def test_cli_dump_list_rng_files_and_tarball(tmp_path):
    """The reference's dump at the save points (src/BayesRRm.cpp:1244-1262 the .lst file, :2805 .rng.<rank> = `file << rng`,
    :2851-2875 `tar -cf <dir>/tarballs/dump_<name>_<it>__<date>.tar -T <out>.lst`)."""
    import tarfile
    d = str(tmp_path)
    write_dataset(d)
    os.makedirs(os.path.join(d, "o", "tarballs"))
    r = subprocess.run(base_args(d, "o", ["--bfile", os.path.join(d, "t")]), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = os.path.join(d, "o", "run")
    lst = open(out + ".lst").read().split("\n")
    assert lst[:4] == [out + ".csv", out + ".xbet", out + ".xcpn", out + ".acu"] and out + ".rng.2" in lst and out + ".mus.0" in lst
    for t in range(3):   # one stream per task: 624 state words and the position, as text
        w = open(out + f".rng.{t}").read().split()
        assert len(w) == 625 and all(x.isdigit() for x in w)
    assert open(out + ".rng.0").read() != open(out + ".rng.1").read()
    tars = os.listdir(os.path.join(d, "o", "tarballs"))
    assert len(tars) == 1 and tars[0].startswith("dump_run_00004__") and tars[0].endswith(".tar")
    names = tarfile.open(os.path.join(d, "o", "tarballs", tars[0])).getnames()
    for ext in (".csv", ".xbet", ".xcpn", ".acu", ".rng.0", ".mrk.2", ".eps.1", ".mus.0"):
        assert any(n.endswith("run" + ext) for n in names), (ext, names)
