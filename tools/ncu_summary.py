"""Dev tool: condenses an `ncu --set full` report (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into the per-kernel
table kept under profiles/, and (optionally) into profiles/ncu_traffic.json, the DRAM bytes per element of the three phase
brackets bench.py reports (`roofline.traffic`).

  python tools/ncu_summary.py gpurun_out/r02f_full_c2.ncu-rep profiles/r02f_ncu_full_c2.csv [--traffic c2 998250]
"""
import csv
import io
import json
import os
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("launch__registers_per_thread", ""),
    ("launch__block_size", ""),
    ("launch__grid_size", ""),
    ("dram__bytes_read.sum", "MB"),
    ("dram__bytes_write.sum", "MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"),
    ("smsp__inst_executed.sum", "inst"),
    ("lts__t_requests_srcunit_tex_op_red.sum", "req"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "req"),
]
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6,
              "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    name_col = idx.get("Kernel Name", idx.get("Function Name"))
    table = []
    for r in data:
        line = {"kernel": r[name_col][:90]}
        for c, want in COLS:
            if c not in idx:
                line[c] = ""
                continue
            v, u = r[idx[c]].replace(",", ""), units[idx[c]]
            try:
                v = float(v)
            except ValueError:
                line[c] = r[idx[c]]
                continue
            if want in ("us", "MB") and u in UNIT_SCALE:
                v *= UNIT_SCALE[u]
            line[c] = v
        table.append(line)
    with open(out, "w") as f:
        f.write("kernel," + ",".join(f"{c} [{u}]" if u else c for c, u in COLS) + "\n")
        for line in table:
            f.write('"' + line["kernel"].replace('"', "'") + '",' + ",".join(f"{line[c]:.6g}" if isinstance(line[c], float) else str(line[c]) for c, _ in COLS) + "\n")
    print(f"{len(table)} launches -> {out}")
    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        workload, n_el = sys.argv[k + 1], float(sys.argv[k + 2])
        bracket = {"element": 0.0, "projection": 0.0, "assembly": 0.0}
        for line in table:
            b = line["dram__bytes_read.sum"] + line["dram__bytes_write.sum"]
            nm = line["kernel"]
            if "second_order" in nm:
                bracket["element"] += b
            elif "project_c_assemble" in nm or "assemble" in nm:
                bracket["assembly"] += b
            elif "project_kernel" in nm:
                bracket["projection"] += b
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
        cur[workload] = {k + "_bytes_per_element": v * 1e6 / n_el for k, v in bracket.items()}
        cur[workload]["source"] = f"dram__bytes_read.sum + dram__bytes_write.sum of the kernels of one evaluation, {os.path.basename(out)} " \
                                  f"(ncu --set full, {int(n_el):,} elements per launch)"
        json.dump(cur, open(path, "w"), indent=1)
        print("traffic ->", path, cur[workload])


if __name__ == "__main__":
    main()
