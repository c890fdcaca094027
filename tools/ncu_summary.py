#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + top stall lines from the source page.
usage: python tools/ncu_summary.py report.ncu-rep [--top 12]"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print("===", d[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:75s} {d[i]:>16s} {units[i]}")
    src = page(rep, "source")
    # several kernels are concatenated: split on the 'Kernel Name' marker rows
    block, name = [], None
    def flush():
        if not block:
            return
        h = block[0]
        si, ci = h.index("# Samples"), h.index("Source")
        body = [r for r in block[1:] if len(r) > si and r[si].replace('.', '').isdigit()]
        tot = sum(float(r[si]) for r in body) or 1.0
        print(f"--- top stall lines: {name[:80]} ({int(tot)} samples)")
        for k, r in sorted(enumerate(body), key=lambda kr: -float(kr[1][si]))[:top]:
            print(f"   {float(r[si]) / tot * 100:5.1f}%  #{k:5d} {r[ci].strip()[:100]}")
    for r in src:
        if r and r[0] == "Kernel Name":
            flush()
            block, name = [], r[1]
        else:
            block.append(r)
    flush()


if __name__ == "__main__":
    main()
