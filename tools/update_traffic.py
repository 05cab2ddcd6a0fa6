"""profiles/pdip_traffic.json from an ncu summary (tools/ncu_summary.py output) of mincurv_pdip_kernel: DRAM bytes per QP stamped
with the SHA-256 of the kernel source the capture was taken from (bench.py reports `traffic` only when the stamp matches).

    python tools/update_traffic.py profiles/r02_pdip_ncu_summary.json [out.json]"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
summ = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "pdip_traffic.json")
d = json.load(open(summ))
src = os.path.join(ROOT, "global_racetrajectory_optimization_b200", "csrc", "mincurv_ipm.cu")
t = {"kernel": "mincurv_pdip_kernel",
     "version": "r02 final (bordered band LDL^T in panels of eight, panel-inverse factor rows, DMMA sweeps, overlapped separator passes)",
     "dram_bytes_per_qp": d["dram_bytes_per_qp"], "n_points": 1000,
     "source": "profiles/r02_pdip_ncu_summary.json (ncu --set full, B=%d launch: dram__bytes_read.sum + dram__bytes_write.sum)" % d["instances"],
     "source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
json.dump(t, open(out, "w"), indent=1)
print(out, t["dram_bytes_per_qp"], t["source_sha256"][:12])
