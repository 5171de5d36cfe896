"""Key metrics of an .ncu-rep (first profiled kernel).
    python scripts/ncu_summary.py rep.ncu-rep                       # print
    python scripts/ncu_summary.py rep.ncu-rep --json KEY FRAMES     # also record the per-frame constants bench.py reads under
                                                                    # profiles/r02_constants.json[KEY] (FRAMES = frames of the launch)"""
import csv, io, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]; v = rows[2] if len(rows) > 2 else rows[1]
val = dict(zip(h, v))
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"{'Kernel Name':95s} {val.get('Kernel Name', '?')}")
for i, k in enumerate(h):
    if k in want or ("issue_stalled" in k and "ratio" in k and float(v[i] or 0) > 0.05):
        print(f"{k:95s} {v[i]}")
if "--json" in sys.argv:
    i = sys.argv.index("--json"); key, frames = sys.argv[i + 1], float(sys.argv[i + 2])
    units = dict(zip(h, rows[1]))
    def num(k, scale_units=True):
        x = float(val[k].replace(",", ""))
        u = units.get(k, "")
        if scale_units:
            x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "msecond": 1.0, "usecond": 1e-3, "second": 1e3, "ms": 1.0, "us": 1e-3, "s": 1e3}.get(u, 1.0)
        return x
    out = ROOT / "profiles" / "r02_constants.json"
    d = json.loads(out.read_text()) if out.exists() else {}
    d[key] = {"kernel": val.get("Kernel Name", "?"), "frames": frames, "duration_ms": num("gpu__time_duration.sum"),
              "warp_instr_per_frame": num("smsp__inst_executed.sum") / frames,
              "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "dram_bytes_per_frame": (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / frames,
              "dram_read_bytes_per_frame": num("dram__bytes_read.sum") / frames, "dram_write_bytes_per_frame": num("dram__bytes_write.sum") / frames,
              "l1tex_data_pipe_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed") if "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed" in val else None,
              "smem_wavefronts_per_frame": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / frames,
              "registers": int(float(val["launch__registers_per_thread"])), "block_size": int(float(val["launch__block_size"])),
              "grid_size": int(float(val["launch__grid_size"])), "capture": Path(rep).name,
              "how": "ncu --set full --clock-control none, one launch (see profiles/README.md)"}
    out.write_text(json.dumps(d, indent=1) + "\n")
    print("recorded", key, "->", out)
