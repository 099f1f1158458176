"""Host-logic tests of the drop-in layer WITHOUT a GPU: the reference's whole provers run through the drop-in's own
objects (libiop_b200/lib/obj_ref/*.o — the shadows under libiop_b200/dropin/ and the adapter under libiop_b200/host/)
linked against tests/hostsim/b200_hostsim.cpp, a host simulator of the C ABI built on the parity oracle.  Everything the
drop-in does on the HOST — which vector goes to which call, residency and lifetime of the device copies, the shadowed
IOP / protocol code, the zk salt stream, transcript assembly — is then checked by transcript equality with the
fingerprints of the unmodified CPU reference (tests/golden/).  The kernels themselves are checked by the -m gpu tests.
The simulator is test infrastructure: see the header of tests/hostsim/b200_hostsim.cpp."""
import contextlib
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "hostsim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libiop_b200_dropin_hostsim.so")


@contextlib.contextmanager
def quiet_stdout():
    """the reference printf()s its parameter blocks"""
    sys.stdout.flush()
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)


@pytest.fixture(scope="module")
def sim():
    if not os.path.isdir(os.path.join(ROOT, "libiop_b200", "lib", "obj_ref")):
        pytest.skip("the drop-in objects were not built here (libiop_b200/dropin/Makefile needs the reference checkout)")
    subprocess.check_call(["make", "-C", SIM_DIR], stdout=subprocess.DEVNULL)
    from libiop_b200 import prover
    return prover.DropinProver(SIM_LIB)


def _golden_cases():
    g = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))
    cases = []
    for line in g["aurora_fingerprints"]:
        log_n, zk, nbytes, sha = str(line).split(",")
        cases.append(("aurora_gf192", int(log_n), zk == "1", int(nbytes), sha))
    for line in g["prover_fingerprints"]:
        which, log_n, zk, nbytes, sha = str(line).split(",")
        cases.append((which, int(log_n), zk == "1", int(nbytes), sha))
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "prover_fingerprints_extra.json"))):
        # 2^14 = the smallest size at which the constraint matrices are served from the device (b200_constraint_matrices.hpp): the
        # pipelined A z / B z / C z extensions (Aurora), the transposed products and the index oracles (Fractal) run their device paths
        if c["log_n"] <= 12 or (c["prover"] in ("aurora_gf192", "aurora_edwards", "fractal_edwards") and c["log_n"] <= 14):
            cases.append((c["prover"], c["log_n"], c["zk"], c["bytes"], c["sha256"]))
    return cases


@pytest.mark.parametrize("which,log_n,zk,nbytes,sha", _golden_cases(),
                         ids=lambda v: str(v) if not isinstance(v, str) or len(v) < 20 else v[:8])
def test_whole_provers_through_the_dropin_host_logic(sim, which, log_n, zk, nbytes, sha):
    with quiet_stdout():
        if which == "aurora_gf192":
            _, dump, acc = sim.aurora_gf192(log_n, make_zk=zk)
        elif which == "aurora_edwards":
            _, dump, acc = sim.aurora_edwards(log_n)
        elif which == "fractal_edwards":
            _, dump, acc, _ = sim.fractal_edwards(log_n)
        else:
            _, dump, acc = sim.ligero_gf192(log_n, make_zk=zk)
    assert acc == 1, "the reference verifier rejected the proof"
    assert (len(dump), hashlib.sha256(dump).hexdigest()) == (nbytes, sha)


AURORA_VARIATIONS = [  # (field, log_n, k_inputs, zk, rs_extra, fri_loc): shapes the golden fingerprints do not cover
    ("gf192", 12, 0, False, 3, 2), ("gf192", 12, 1, False, 3, 2), ("gf192", 12, 7, True, 3, 2), ("gf192", 12, 31, False, 2, 1),
    ("gf192", 12, 3, True, 2, 3), ("gf192", 13, 15, False, 3, 3), ("gf192", 12, 255, False, 3, 2),
    ("edwards", 12, 0, False, 3, 2), ("edwards", 12, 3, False, 3, 2), ("edwards", 12, 31, False, 2, 1), ("edwards", 13, 7, False, 3, 3)]


@pytest.mark.parametrize("field,log_n,k_inputs,zk,rs_extra,fri_loc", AURORA_VARIATIONS)
def test_aurora_parameter_variations_against_the_cpu_reference(sim, field, log_n, k_inputs, zk, rs_extra, fri_loc):
    """The evaluation-domain forms of the prover's coefficient chains (witness codeword, sumcheck quotient, lincheck
    polynomials: DESIGN.md 3.8b) at other input counts, rates and FRI localizations than the committed fingerprints: the
    unmodified CPU reference compiled here (oracle/_ref) and the drop-in over the simulator must emit the same bytes."""
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libiopref.so not built (needs the reference checkout)")
    ref = Reference()
    with quiet_stdout():
        if field == "gf192":
            want = ref.aurora_gf192(log_n, k_inputs=k_inputs, make_zk=zk, rs_extra=rs_extra, fri_loc=fri_loc)
            got = sim.aurora_gf192(log_n, k_inputs=k_inputs, make_zk=zk, rs_extra=rs_extra, fri_loc=fri_loc)
        else:
            want = ref.aurora_edwards(log_n, k_inputs=k_inputs, rs_extra=rs_extra, fri_loc=fri_loc)
            got = sim.aurora_edwards(log_n, k_inputs=k_inputs, rs_extra=rs_extra, fri_loc=fri_loc)
    assert got[2] == 1, "the reference verifier rejected the proof"
    assert bytes(got[1]) == bytes(want[1])


@pytest.mark.parametrize("log_n,num_variables,k_inputs,zk", [(12, 8191, 31, False), (12, 2047, 7, False), (13, 4095, 3, True),
                                                             (14, 32767, 31, False), (14, 8191, 31, False)])
def test_aurora_with_unequal_constraint_and_variable_domains(sim, log_n, num_variables, k_inputs, zk):
    """Constraint systems whose variable domain is larger or smaller than the constraint domain (the summation domain is the
    larger of the two: placement strides, sizes of z and of the matrices' columns all differ from the square case), against
    the unmodified CPU reference compiled here; at 2^14 constraints the device-matrix paths run."""
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libiopref.so not built (needs the reference checkout)")
    ref = Reference()
    before = sim.lazy_stats()["faults"]
    with quiet_stdout():
        want = ref.aurora_gf192_shape(log_n, num_variables, k_inputs=k_inputs, make_zk=zk)
        got = sim.aurora_gf192_shape(log_n, num_variables, k_inputs=k_inputs, make_zk=zk)
    assert want[2] == 1 and got[2] == 1
    assert bytes(got[1]) == bytes(want[1])
    assert sim.lazy_stats()["faults"] == before


def test_ligero_and_fractal_parameter_variations_against_the_cpu_reference(sim):
    """Other height/width ratios, rates and FRI localizations than the committed fingerprints (Ligero incl. zero knowledge)."""
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libiopref.so not built (needs the reference checkout)")
    ref = Reference()
    with quiet_stdout():
        pairs = [(ref.ligero_gf192(13, ratio=0.05, rs_extra=3), sim.ligero_gf192(13, ratio=0.05, rs_extra=3)),
                 (ref.ligero_gf192(12, ratio=0.2, rs_extra=2, make_zk=True), sim.ligero_gf192(12, ratio=0.2, rs_extra=2, make_zk=True)),
                 (ref.ligero_gf192(12, ratio=1.0, rs_extra=2), sim.ligero_gf192(12, ratio=1.0, rs_extra=2)),
                 (ref.fractal_edwards(12, rs_extra=2, fri_loc=1)[:3], sim.fractal_edwards(12, rs_extra=2, fri_loc=1)[:3]),
                 (ref.fractal_edwards(13, rs_extra=3, fri_loc=3)[:3], sim.fractal_edwards(13, rs_extra=3, fri_loc=3)[:3])]
    for want, got in pairs:
        assert got[2] == 1 and bytes(got[1]) == bytes(want[1])


def test_device_matrix_paths_at_other_parameters_against_the_cpu_reference(sim):
    """N = 2^14 is the smallest size at which the constraint matrices are served from the device (pipelined A z / B z / C z
    extensions, transposed products, index oracles): Aurora with 3 inputs at rate 1/4, localization 3, and Fractal at rate 1/4,
    localization 1 — neither among the committed fingerprints — against the unmodified CPU reference compiled here."""
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libiopref.so not built (needs the reference checkout)")
    ref = Reference()
    with quiet_stdout():
        want_a = ref.aurora_gf192(14, k_inputs=3, make_zk=False, rs_extra=2, fri_loc=3)
        got_a = sim.aurora_gf192(14, k_inputs=3, make_zk=False, rs_extra=2, fri_loc=3)
        want_f = ref.fractal_edwards(14, rs_extra=2, fri_loc=1)
        got_f = sim.fractal_edwards(14, rs_extra=2, fri_loc=1)
    assert got_a[2] == 1 and bytes(got_a[1]) == bytes(want_a[1])
    assert got_f[2] == 1 and bytes(got_f[1]) == bytes(want_f[1])


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_matrix_flattening_on_host_workers(sim, threads):
    """The R1CS matrices are flattened by a few persistent host threads (b200_constraint_matrices.hpp: host_workers).  The
    split of the rows over the workers must not show in the result: the same Aurora N = 2^14 proof (the smallest size served
    from the device matrices) with 1, 3 and 8 workers, twice each, against the committed fingerprint."""
    want = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "prover_fingerprints_extra.json")))
            if c["prover"] == "aurora_gf192" and c["log_n"] == 14 and not c["zk"]][0]
    saved = os.environ.get("B200_HOST_THREADS")
    os.environ["B200_HOST_THREADS"] = threads
    try:
        for _ in range(2):
            with quiet_stdout():
                _, dump, acc = sim.aurora_gf192(14, make_zk=False)
            assert acc == 1 and (len(dump), hashlib.sha256(dump).hexdigest()) == (want["bytes"], want["sha256"])
    finally:
        if saved is None:
            os.environ.pop("B200_HOST_THREADS", None)
        else:
            os.environ["B200_HOST_THREADS"] = saved


def test_provers_do_not_read_device_resident_vectors_on_the_host(sim):
    """Oracles the adapter produces stay on the device behind lazy host vectors; a host read is served correctly by the fault
    handler but costs a whole-vector download.  None of the provers may take that path (the same check runs on the GPU box in
    test_gpu_parity.py; here it catches a shadow that starts touching a lazy vector before any GPU time is spent)."""
    before = sim.lazy_stats()
    with quiet_stdout():
        sim.aurora_gf192(12, make_zk=False, verify=False)
        sim.aurora_gf192(14, make_zk=False, verify=False)
        sim.aurora_gf192(12, make_zk=True, verify=False)
        sim.aurora_edwards(12, verify=False)
        sim.fractal_edwards(12, verify=False)
        sim.ligero_gf192(12, make_zk=False, verify=False)
    after = sim.lazy_stats()
    assert after["gathers"] > before["gathers"]
    assert after["faults"] == before["faults"], (before, after)


def test_lazy_vectors_against_the_simulator():
    """tests/hostsim/lazy_vector_test.cpp: the device-resident ("lazy") host vectors of libiop_b200/host/b200_residency.hpp —
    registration by ownership, materialisation on any host access (read, write, by-value copy) through the SIGSEGV handler,
    release on destruction, device-to-device clones, query gathers.  Needs only the simulator, not the reference checkout."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", SIM_DIR, "sim"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(SIM_DIR, "_build", "lazy_vector_test")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "0 failed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


def test_simulator_is_not_reachable_from_the_product():
    """nothing under libiop_b200/ or include/ names the simulator or the oracle, and the product drop-in library does not link them"""
    for base in ("libiop_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc", ".log")) or os.path.basename(d) in ("obj", "obj_ref", "__pycache__"):
                    continue
                text = open(os.path.join(d, f), errors="ignore").read()
                assert "hostsim" not in text or f == "prover.py", os.path.join(d, f)     # prover.py only documents the lib_path argument
                assert "liboracle" not in text, os.path.join(d, f)
    lib = os.path.join(ROOT, "libiop_b200", "lib", "libiop_b200_dropin.so")
    if os.path.exists(lib):
        needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
        assert "hostsim" not in needed and "liboracle" not in needed
