"""Summarise an .ncu-rep (or a --csv metrics log) into the few numbers DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import io
import re
import subprocess
import sys

PAT = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed|lts__t_sector_hit_rate\.pct|"
    r"sm__cycles_elapsed\.avg\.per_second|launch__registers_per_thread|launch__grid_size|launch__block_size|"
    r"launch__cluster_size|l1tex__m_xbar2l1tex_read_bytes\.sum|lts__t_bytes\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor\.sum|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|smsp__inst_executed\.sum|launch__shared_mem_per_block_dynamic)$")


def main(path):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("kernel:", d.get("Kernel Name", "?")[:150])
            for h, u in zip(hdr, units):
                if PAT.match(h):
                    print(f"  {h:80s} {d[h]:>24s} {u}")
    else:
        lines = [l for l in open(path) if not l.startswith("==")]
        for r in csv.DictReader(io.StringIO("".join(lines))):
            print(f"{r['Kernel Name'][:70]:72s} {r['Metric Name']:75s} {r['Metric Value']:>22s} {r['Metric Unit']}")


if __name__ == "__main__":
    main(sys.argv[1])
