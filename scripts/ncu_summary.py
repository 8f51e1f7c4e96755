"""Turn an .ncu-rep into the small JSON summaries kept under profiles/.

    python scripts/ncu_summary.py full  gpurun_out/x.ncu-rep profiles/r01_ncu_X_full.json "command line that made it"
    python scripts/ncu_summary.py list  gpurun_out/launches.csv profiles/r01_launch_list_summary.json "command line"
"""
import csv, io, json, subprocess, sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum"]


def full(rep, out, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = [{k: r[ix[k]] for k in KEEP if k in ix} for r in rows[2:]]
    json.dump({"source": cmd, "units": {k: units[ix[k]] for k in KEEP if k in ix}, "launches": launches}, open(out, "w"), indent=1)
    print(out, len(launches), "launches")


def launch_list(path, out, cmd):
    rows = [r for r in csv.reader(open(path)) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] in ("ns", "nsecond") else (v * 1e3 if r[mu] in ("ms", "msecond") else v)   # -> us
        name = r[kn].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    total = sum(a[1] for a in agg.values())
    ks = [{"kernel": k, "launches": a[0], "total_ms": round(a[1] / 1e3, 3), "share": round(a[1] / total, 4),
           "avg_us": round(a[1] / a[0], 2)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump({"source": cmd, "launches": sum(a[0] for a in agg.values()), "total_ms": round(total / 1e3, 3), "kernels": ks},
              open(out, "w"), indent=1)
    print(out, ks[:4])


if __name__ == "__main__":
    (full if sys.argv[1] == "full" else launch_list)(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
