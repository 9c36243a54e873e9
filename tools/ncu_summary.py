"""Summarise `ncu --set full` reports: python tools/ncu_summary.py report.ncu-rep [...]  (reads them with `ncu -i ... --page raw --csv`).
One line per captured launch: duration, DRAM bytes read / written, DRAM throughput % of ncu's peak, tensor-pipe active %,
shared-memory wavefront %, registers per thread, SM clock."""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.avg.per_second", "clk")]
SCALE = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def main():
    print("# kernel | duration ms | dram read GB | dram write GB | dram TB/s | tensor pipe %active | smem wavefronts %peak | regs/thread | SM clock GHz")
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        name_i = idx["Kernel Name"]

        def col(metric):
            for h, i in idx.items():
                if h.endswith(metric):
                    return i
            return None
        print(f"## {rep}")
        for r in data:
            out = []
            for metric, _ in COLS:
                i = col(metric)
                try:
                    v = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0) if i is not None else None
                except ValueError:
                    v = None
                if v is None:
                    out.append("n/a")
                    continue
                out.append(f"{v:.3f}" if v < 100 else f"{v:.0f}")
            try:
                out.insert(3, f"{(float(out[1]) + float(out[2])) / float(out[0]):.2f}")
            except ValueError:
                out.insert(3, "n/a")
            print(f"{r[name_i][:40]:40s} | " + " | ".join(out))


if __name__ == "__main__":
    main()
