"""SASS evidence table: which kernels of libps_b200.so carry tcgen05 / TMA / TMEM / bulk-copy / mbarrier / PDL / reduction / system-scope
instructions.  CPU only:  python scripts/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ps_b200", "lib", "libps_b200.so")
COLS = [("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("UTCATOMSWS", r"\bUTCATOMSWS"),
        ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("PREEXIT", r"\bPREEXIT"), ("ACQBULK", r"\bACQBULK"),
        ("REDG", r"\bREDG?\.E"), ("ATOMS", r"\bATOMS"), ("MATCH", r"\bMATCH"), ("LD.SYS", r"\bLDG\.E[.\w]*\.STRONG\.SYS"),
        ("ST.SYS", r"\bSTG\.E[.\w]*\.STRONG\.SYS"), ("MEMBAR.SYS", r"\bMEMBAR\.\w+\.SYS"), ("MEMBAR.GPU", r"\bMEMBAR\.\w+\.GPU"), ("MUFU", r"\bMUFU")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            mangled = m.group(1)
            d = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
            base = re.sub(r"^void ", "", d).replace("(anonymous namespace)::", "")
            base = re.sub(r"<.*", "", base.split("(")[0]).split("::")[-1]
            name = base
            per.setdefault(name, []).append(collections.Counter())
            continue
        if name is None:
            continue
        for col, pat in COLS:
            if re.search(pat, line):
                per[name][-1][col] += 1
    print("# SASS evidence, round 2 — `cuobjdump -sass ps_b200/lib/libps_b200.so` (sm_100a only; `scripts/sass_evidence.py`)\n")
    print("Maximum count of each mnemonic over the template instantiations of a kernel. `UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA tensor load, `LDTM` = tcgen05.ld,")
    print("`UTCBAR` = tcgen05.commit, `UTCATOMSWS` = TMEM alloc / dealloc, `UBLKCP` = cp.async.bulk (TMA bulk copy: hot rows of the lookup, records of the update),")
    print("`LDGSTS` = cp.async 16 B (row and delta staging), `SYNCS` = mbarrier, `PREEXIT` / `ACQBULK` = griddepcontrol.launch_dependents / .wait, `REDG` = red.global,")
    print("`ATOMS` = shared-memory atomics, `MATCH` = match.any, `LD.SYS` / `ST.SYS` = system-scope loads / stores (peer memory, flags), `MEMBAR.SYS` / `.GPU` = fences")
    print("(producer blocks fence at GPU scope, the publishing block at system scope), `MUFU` = approximate root / reciprocal.\n")
    print("| kernel | inst. | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for k, insts in per.items():
        mx = {c: max(i[c] for i in insts) for c, _ in COLS}
        if not any(mx.values()):
            continue
        print(f"| `{k}` | {len(insts)} | " + " | ".join(str(mx[c]) for c, _ in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
