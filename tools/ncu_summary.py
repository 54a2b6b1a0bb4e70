"""Compact per-kernel table from `ncu -i X.ncu-rep --page raw --csv` (one line per captured launch)."""
import csv, re, subprocess, sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("# " + path)
    print("# kernel | " + " | ".join(n for _, n in COLS))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")       # drop the parameter list
        head = name.split("<", 1)[0]
        name = name[head.rfind("::") + 2:] if "::" in head else name    # drop namespaces in front of the kernel name
        cells = []
        for key, _ in COLS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("-")
        print(f"{name:44s} | " + " | ".join(cells))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
