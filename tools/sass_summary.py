"""Per-kernel opcode counts of libgecco_b200.so (cuobjdump -sass): the SASS evidence that the hot kernels run on the
Blackwell tensor / TMA / TMEM paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA ->
UTMALDG/UTMASTG, legacy mma.sync -> HMMA).  Writes profiles/<tag>_sass_opcodes.txt.

    python tools/sass_summary.py r2
"""
import collections, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "gecco_b200" / "lib" / "libgecco_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "MUFU.EX2",
         "FFMA2", "FENCE.VIEW.ASYNC", "ELECT", "LDS", "STS", "LDG", "STG", "ATOMG", "REDG", "RED", "DADD", "BAR"]


def main(tag: str):
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("gecco::", "")
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    lines = [f"SASS opcode counts per kernel of {LIB.name} (cuobjdump -sass, sm_100a); columns with a zero everywhere are omitted.",
             "UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store, "
             "UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (fallback attention kernels only).", ""]
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    lines.append(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
    for name, c in kernels.items():
        lines.append(f"{name[:78]:78s} {c['_total']:7d} " + " ".join(f"{c[w]:9d}" for w in cols))
    dst = ROOT / "profiles" / f"{tag}_sass_opcodes.txt"
    dst.write_text("\n".join(lines) + "\n")
    print(dst, len(kernels), "kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
