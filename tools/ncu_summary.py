"""Summarises .ncu-rep files (read here, no GPU): key raw metrics + the hottest SASS lines by stall samples.
Usage: python tools/ncu_summary.py rep1.ncu-rep [rep2 ...] [--top N]"""
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
           "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
           "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_uniform.sum"]
def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout
def main():
    top = 14
    reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
    if "--top" in sys.argv: top = int(sys.argv[sys.argv.index("--top") + 1])
    for rep in reps:
        rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("== %s | %s grid %s block %s" % (rep, r[4][:50], r[8], r[7]))
            for m in METRICS:
                if m in hdr: print("   %-66s %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
        h = None; body = []
        for r in src:
            if r and r[0] == "Address": h = r; continue
            if h and len(r) > 5: body.append(r)
        if not h: continue
        si, ie = h.index("# Samples"), h.index("Instructions Executed")
        tot = sum(int(r[si]) for r in body) or 1
        print("   total stall samples %d; hottest SASS lines:" % tot)
        for i, r in sorted(sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:top]):
            print("   %5d %-64s %6s (%4.1f%%) exec %s" % (i, r[1].strip()[:64], r[si], 100.0 * int(r[si]) / tot, r[ie]))
if __name__ == "__main__":
    main()
