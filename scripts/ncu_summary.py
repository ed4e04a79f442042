"""Print the handful of ncu metrics the roofline argument needs from a .ncu-rep (runs without a GPU)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_bytes.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def main(path, json_out=None, kernel_filter=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d['Kernel Name'][:110]}  grid {d['Grid Size']} block {d['Block Size']}")
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    print(f"   {k:82s} {d[h]:>16s} {u[h]}")
                    break
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            tu = u["gpu__time_duration.sum"]
            t_us = t * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(tu, 1)
            def b(k):
                v = float(d[k].replace(",", ""))
                return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u[k], 1)
            tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            print(f"   dram traffic {tr/1e6:.1f} MB -> {tr/t_us/1e3:.1f} GB/s over {t_us:.1f} us")
            if json_out and (kernel_filter is None or kernel_filter in d["Kernel Name"]):
                import json

                with open(json_out, "w") as f:  # the first matching launch: what bench.py reads as roofline.traffic
                    json.dump({"kernel": d["Kernel Name"][:160], "grid": d["Grid Size"], "block": d["Block Size"],
                               "duration_us_under_ncu": t_us, "dram_bytes_read": b("dram__bytes_read.sum"),
                               "dram_bytes_write": b("dram__bytes_write.sum"), "source_report": path,
                               "how": "ncu --set full --clock-control none, one launch"}, f, indent=1)
                json_out = None
        except Exception as e:  # noqa
            print("   (traffic n/a)", e)


if __name__ == "__main__":
    # ncu_summary.py report.ncu-rep [--json out.json [--kernel substring]]
    a = sys.argv[1:]
    jo = a[a.index("--json") + 1] if "--json" in a else None
    kf = a[a.index("--kernel") + 1] if "--kernel" in a else None
    main(a[0], jo, kf)
