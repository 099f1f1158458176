"""Static SASS instruction counts per kernel of a built library / binary (cuobjdump -sass): the evidence table of
profiles/r2_sass_summary.md.  Usage: python bench/sass_summary.py <file> [<file> ...]  (prints markdown tables)"""
import collections, re, subprocess, sys

COLS = [("LDG.128", r"\bLDG\.E\.(?:\w+\.)*128"), ("LDG.64", r"\bLDG\.E\.(?:\w+\.)*64"), ("STG.128", r"\bSTG\.E\.(?:\w+\.)*128"),
        ("STG.64", r"\bSTG\.E\.(?:\w+\.)*64"), ("LDS.128", r"\bLDS\.128"), ("STS.128", r"\bSTS\.128"), ("SHFL", r"\bSHFL\."),
        ("UBLKCP", r"\bUBLKCP"), ("IMAD.WIDE", r"\bIMAD\.WIDE"), ("LOP3", r"\bLOP3"), ("BAR", r"\bBAR\."), ("ATOM/RED", r"\b(?:ATOM|ATOMG|ATOMS|RED)\b")]


def summarise(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    demangled = subprocess.run(["cu++filt"], input=out, capture_output=True, text=True).stdout or out
    rows, cur, counts, n = [], None, None, 0
    for line in demangled.splitlines():
        m = re.match(r"\s*Function : (.*)", line)
        if m:
            if cur: rows.append((cur, n, counts))
            cur = re.sub(r"\([^()]*\)\s*$", "", m.group(1)).strip()            # drop the parameter list, keep the template arguments
            cur = cur.replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
            counts = collections.Counter(); n = 0
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            n += 1
            for name, pat in COLS:
                if re.search(pat, line): counts[name] += 1
    if cur: rows.append((cur, n, counts))
    print(f"## {path}\n")
    print("| kernel | instr | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for name, n, c in rows:
        print(f"| `{name}` | {n} | " + " | ".join(str(c[k]) for k, _ in COLS) + " |")
    print()


if __name__ == "__main__":
    for f in sys.argv[1:]:
        summarise(f)
