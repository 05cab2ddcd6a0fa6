"""Condenses an .ncu-rep (ncu --set full, one kernel launch) into the JSON summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_pdip_v6.ncu-rep profiles/r01_v6_pdip_ncu_summary.json [B]

B = instances in the profiled launch (adds dram bytes per QP).  Reads the raw and the source (SASS) pages."""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    B = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    col = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    summ = {"kernel": col.get("Kernel Name", ("?",))[0], "metrics": {k: list(col[k]) for k in KEYS if k in col}}
    stalls = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(v)
              for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h}
    summ["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
    if B:
        def gb(k):
            v, u = col[k]
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        summ["instances"] = B
        summ["dram_bytes_per_qp"] = (gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")) / B
    src = page(rep, "source", ("--print-source", "sass"))
    h2 = src[1]
    i_src, i_s, i_ex = h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
    rows = [(int(r[i_s] or 0), r[i_src].strip(), int(r[i_ex] or 0)) for r in src[2:] if len(r) > i_s and (r[i_s] or "0").isdigit()]
    tot = sum(r[0] for r in rows) or 1
    summ["sass_instructions"] = len(rows)
    summ["top_sampled_sass"] = [{"pct_samples": round(100.0 * s / tot, 2), "executed": e, "sass": t} for s, t, e in sorted(rows, reverse=True)[:12]]
    mix = {}
    for s, t, e in rows:
        op = t.split()[1].split(".")[0] if t.startswith("@") and len(t.split()) > 1 else t.split(".")[0].split()[0] if t else "?"
        mix[op] = mix.get(op, 0) + e
    summ["executed_instruction_mix_top"] = dict(sorted(mix.items(), key=lambda kv: -kv[1])[:14])
    json.dump(summ, open(dst, "w"), indent=1)
    print(json.dumps({k: summ[k] for k in ("kernel", "warp_stalls_per_issue")} | {"dram_bytes_per_qp": summ.get("dram_bytes_per_qp")}, indent=1)[:1500])


if __name__ == "__main__":
    main()
