"""One line per kernel launch of an ncu report with the metrics that decide an HBM-bound kernel.  Usage: python tools/ncu_table.py <rep>"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdGB"), ("dram__bytes_write.sum", "wrGB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ldsect"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ldreq"),
        ("smsp__inst_executed.sum", "inst")]
units = rows[1]
print("kernel".ljust(64), " ".join(n.rjust(8) for _, n in want), " top stalls")
for r in rows[2:]:
    name = r[col["Kernel Name"]][:64]
    out = []
    for m, n in want:
        if m not in col:
            out.append("-".rjust(8)); continue
        v = r[col[m]].replace(",", "")
        try:
            x = float(v)
            u = units[col[m]]
            if n == "us":
                x = x / 1e3 if u == "ns" else (x * 1e3 if u == "ms" else x)
            if n in ("rdGB", "wrGB"):
                x = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(u, 1.0) * x
            out.append((f"{x:8.2f}" if x < 1e5 else f"{x:8.3g}"))
        except ValueError:
            out.append(v[:8].rjust(8))
    stalls = []
    for h in hdr:
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                stalls.append((float(r[col[h]].replace(",", "")), h.split("stalled_")[1]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    tot = sum(x for x, _ in stalls) or 1.0
    print(name.ljust(64), " ".join(out), " ", ", ".join(f"{n} {100 * x / tot:.0f}%" for x, n in stalls[:4]))
