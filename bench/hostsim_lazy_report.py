"""Runs the provers through the host simulator (tests/hostsim) and reports how many device-resident vectors the HOST touched
(each touch = a whole-vector download on the real device) and the bytes that crossed the simulated PCIe.  With
B200_LAZY_TRACE=1 in the environment every touch prints a backtrace (resolve with addr2line -e <hostsim .so>).
Usage: python bench/hostsim_lazy_report.py [aurora|fractal|ligero|aurora_zk|aurora_edwards] [log_n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("B200_DETERMINISTIC_RNG", "1")
from libiop_b200 import prover
from test_dropin_hostsim import quiet_stdout, SIM_LIB

which = sys.argv[1] if len(sys.argv) > 1 else "aurora"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
p = prover.DropinProver(SIM_LIB)
d0, l0 = p.device_stats(), p.lazy_stats()
with quiet_stdout():
    if which == "aurora": r = p.aurora_gf192(log_n, verify=False)
    elif which == "aurora_zk": r = p.aurora_gf192(log_n, make_zk=True, verify=False)
    elif which == "aurora_edwards": r = p.aurora_edwards(log_n, verify=False)
    elif which == "fractal": r = p.fractal_edwards(log_n, verify=False)
    else: r = p.ligero_gf192(log_n, verify=False)
d1, l1 = p.device_stats(), p.lazy_stats()
print(which, log_n, "host touches:", l1["faults"] - l0["faults"], "bytes:", l1["fault_bytes"] - l0["fault_bytes"], "query gathers:", l1["gathers"] - l0["gathers"],
      "| h2d", d1["h2d_bytes"] - d0["h2d_bytes"], "d2h", d1["d2h_bytes"] - d0["d2h_bytes"], "| codeword bytes", 24 << (log_n + 3), p.residency_stats())
