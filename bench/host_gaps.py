"""Where the host spends the time BETWEEN C-ABI calls during one Aurora proof through the drop-in: runs the prover with
B200_TRACE=1 (every call synchronised and time-stamped on stderr) and lists the calls of the last proof with the host gap
before each.  Usage: python bench/host_gaps.py [log_n] [min_gap_ms] [raw_out|-] [aurora|aurora_zk|fractal|ligero]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
min_gap = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw_out = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None      # also write the raw merged stream (API calls + enter/leave blocks) of the last proof
which = sys.argv[4] if len(sys.argv) > 4 else "aurora"
call = {"aurora": "p.aurora_gf192_phases(%d,1)" % log_n, "aurora_zk": "p.aurora_gf192(%d, make_zk=True, verify=False)[0]" % log_n,
        "fractal": "p.fractal_edwards(%d, verify=False)[0]" % log_n, "ligero": "p.ligero_gf192(%d, verify=False)[0]" % log_n}[which]
code = ("import sys,os,ctypes as C; sys.path.insert(0,%r); from libiop_b200.prover import DropinProver; p=DropinProver();\n"
        "libc=C.CDLL(None); libc.setvbuf(C.c_void_p.in_dll(libc,'stdout'), None, 2, 0); os.dup2(2,1)\n"
        "for r in range(3):\n    sys.stderr.write('=== proof %%d\\n' %% r); sys.stderr.flush(); s=%s; sys.stderr.write('=== phases %%s\\n' %% s); sys.stderr.flush()\n" % (ROOT, call))
r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B200_TRACE="1", B200_DETERMINISTIC_RNG="1", **({"B200_REF_PROFILE": "1"} if raw_out else {})),
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
lines = r.stderr.splitlines()
start = max(i for i, l in enumerate(lines) if l.startswith("=== proof"))
if raw_out:
    with open(raw_out, "w") as f:
        for l in lines[start:]:
            if l.startswith("b200 ") or "(enter)" in l or "(leave)" in l or l.startswith("==="): f.write(l[:150] + "\n")
calls = []
for l in lines[start:]:
    m = re.match(r"b200 (\S+)\s+(\d+)\s+([\d.]+) ms\s+@\s*([\d.]+)", l)
    if m: calls.append((m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4))))
    elif l.startswith("=== phases"): print(l)
print(f"{len(calls)} calls; device-call time {sum(c[2] for c in calls):.1f} ms; span {calls[-1][3] + calls[-1][2] - calls[0][3]:.1f} ms")
prev_end = calls[0][3]
print("    at_ms   gap_before  call                          detail        ms")
tot_gap = 0.0
for name, detail, ms, at in calls:
    gap = at - prev_end; tot_gap += max(gap, 0)
    if gap >= min_gap or ms >= 2.0:
        print(f"{at - calls[0][3]:9.2f} {gap:10.2f}   {name:28s} {detail:12d} {ms:9.3f}")
    prev_end = at + ms
print(f"host time between calls: {tot_gap:.1f} ms")
