"""Per-eval counts of the k_cost main launch from an `ncu --set full --import-source on` capture, for bench.py's roofs.
usage: python scripts/ncu_counts.py gpurun_out/X.ncu-rep EVALS_OF_THE_LAUNCH out.json ["source note"]
Writes {warp_inst_per_eval, fp64_warp_inst_per_eval, l2_tex_read_sectors_per_eval, dram_bytes_per_launch, ...}: warp instructions
from smsp__inst_executed.sum, fp64-pipe warp instructions counted from the SASS page (opcodes D*), L2 sectors from
lts__t_sectors_srcunit_tex_op_read.sum, DRAM bytes from dram__bytes_read.sum + dram__bytes_write.sum."""
import csv, io, json, subprocess, sys

rep, evals, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
gi = h.index("launch__grid_size")
cost = [r for r in data if "k_cost" in r[h.index("Kernel Name")]]
idx_all = [i for i, r in enumerate(data) if "k_cost" in r[h.index("Kernel Name")]]
main = max(cost, key=lambda r: float(r[gi]))
main_pos = [i for i in idx_all if data[i] is main][0]


def val(name):
    i = h.index(name)
    v, u = float(main[i]), units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    return v * scale


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Address":
        cur = {"hdr": r, "rows": []}
        kern.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
def tally(k):
    iS, iI = k["hdr"].index("Source"), k["hdr"].index("Instructions Executed")
    fp64 = tot = 0
    for r in k["rows"]:
        try:
            n = int(r[iI])
        except ValueError:
            continue
        tot += n
        op = r[iS].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
        if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")):
            fp64 += n
    return tot, fp64


# the report lists one or more SASS sections per launch: take the one whose instruction total is the launch's own
want = val("smsp__inst_executed.sum")
tot, fp64 = min((tally(k) for k in kern), key=lambda tf: abs(tf[0] - want))
res = {"kernel": main[h.index("Kernel Name")], "grid": int(float(main[gi])), "evals_of_launch": evals,
       "duration_us_under_ncu": val("gpu__time_duration.sum") / 1e3 if units[h.index("gpu__time_duration.sum")] in ("nsecond", "ns") else val("gpu__time_duration.sum"),
       "warp_inst": val("smsp__inst_executed.sum"), "warp_inst_per_eval": val("smsp__inst_executed.sum") / evals,
       "fp64_warp_inst": fp64, "fp64_warp_inst_per_eval": fp64 / evals, "sass_page_warp_inst": tot,
       "l2_tex_read_sectors": val("lts__t_sectors_srcunit_tex_op_read.sum"),
       "l2_tex_read_sectors_per_eval": val("lts__t_sectors_srcunit_tex_op_read.sum") / evals,
       "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
       "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
       "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
       "source": (note + " " if note else "") + "ncu --set full capture " + rep.split("/")[-1] + " (scripts/ncu_counts.py)"}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
