"""DRAM traffic per launch of each hot kernel from ncu --set full reports -> profiles/<tag>_traffic.json
   python tools/ncu_traffic.py profiles/r01_traffic.json gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep ..."""
import csv, json, subprocess, sys
out, reps = sys.argv[1], sys.argv[2:]
res = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    g = lambda r, k: float(r[hdr.index(k)].replace(",", ""))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].strip()
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd = g(r, "dram__bytes_read.sum") * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = g(r, "dram__bytes_write.sum") * scale[units[hdr.index("dram__bytes_write.sum")]]
        res[name] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                     "duration_us_under_ncu": g(r, "gpu__time_duration.sum"),
                     "tensor_pipe_active_pct": g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "lsu_data_pipe_pct": g(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                     "source": rep.split("/")[-1]}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
