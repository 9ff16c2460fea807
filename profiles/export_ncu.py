"""Export the per-launch metrics the design decisions rest on from an .ncu-rep (run where ncu is installed):
python profiles/export_ncu.py gpurun_out/x.ncu-rep out.csv ["comment line"]"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__cluster_size"]
idx = {w: hdr.index(w) for w in want if w in hdr}
kn = hdr.index("Kernel Name")
with open(out, "w") as f:
    if note:
        f.write("# " + note + "\n")
    f.write("launch,kernel," + ",".join(f"{w} [{units[i]}]" for w, i in idx.items()) + "\n")
    for n, r in enumerate(data):
        name = r[kn].split("(")[0].split("::")[-1].replace(", ", ";").replace(",", ";")
        f.write(f"{n},{name}," + ",".join(r[i] for i in idx.values()) + "\n")
print(open(out).read()[:200])
