"""Turns ncu CSV exports into the small summaries kept under profiles/.

    python tools/ncu_summarise.py launches <launches.csv> <summary.txt> ["header text"]
        from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>`
    python tools/ncu_summarise.py full <raw.csv> <summary.json>
        from `ncu -i report.ncu-rep --page raw --csv > raw.csv` of an `ncu --set full` capture
"""
import collections, csv, json, re, sys

KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("gecco::<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def launches(src, dst, header=""):
    rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mi, ui = hdr.index("Metric Name"), hdr.index("Metric Unit")
    acc = collections.OrderedDict()
    total, n = 0.0, 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = acc.setdefault(r[ki], [0.0, 0])
        a[0] += us
        a[1] += 1
        total += us
        n += 1
    with open(dst, "w") as f:
        if header:
            f.write(header.rstrip() + "\n")
        f.write(f"# total {total / 1e3:.2f} ms over {n} launches\n")
        for k, (us, c) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  n={c:4d} avg={us / c:8.1f}us  {k[:100]}\n")


def full(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for i, r in enumerate(data):
        d = dict(zip(hdr, r))
        e = {"launch_index": i, "kernel": short(d.get("Kernel Name", ""))}
        for k in KEEP:
            if k in d:
                e[k] = f"{d[k]} {units[hdr.index(k)]}".strip()
        out.append(e)
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
