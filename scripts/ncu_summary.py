"""Summarise ncu --set full captures for profiles/: python scripts/ncu_summary.py rep1.ncu-rep [rep2 ...]
Per captured launch: duration, instruction count, issue / pipe utilisation, occupancy, DRAM bytes, stall reasons; for k_gn also the
opcode mix per (warp, source point) from the source page."""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


for rep in sys.argv[1:]:
    rows = page(rep, "raw")
    h, units = rows[0], rows[1]
    print(f"== {rep}")
    for r in rows[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        name = re.sub(r"\(.*", "", d.get("Kernel Name", "?"))[:40]
        print(f"  kernel={name} | " + " | ".join(f"{k.split('.')[0]}={d.get(k)}{(' ' + u.get(k, '')) if 'bytes' in k or 'time' in k else ''}" for k in KEYS if k in d))
        st = {k: float(v.replace(",", "")) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("_not_issued") and v not in ("", "n/a")}
        tot = sum(st.values()) or 1.0
        print("    stall reasons (share of samples): " + ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
    # opcode mix of the k_gn launch(es) from the source page
    src = page(rep, "source")
    cur, ops, total = None, collections.Counter(), 0
    hdr = None
    for r in src:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and cur and "k_gn" in cur and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", d["Source"])
            if m:
                n = int(d["Instructions Executed"] or 0)
                ops[m.group(1)] += n
                total += n
    if total:
        print(f"    k_gn source page: {total} warp instructions; top opcodes (share): " + ", ".join(f"{o} {100 * c / total:.1f}%" for o, c in ops.most_common(16)))
