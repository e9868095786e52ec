"""Writes profiles/<tag>_ncu_traffic.json (DRAM bytes per launch of the dominant kernel) from an `ncu --set full` capture:
   python tools/ncu_traffic.py gpurun_out/x.ncu-rep profiles/r02_ncu_traffic.json <rows_per_launch>"""
import csv, io, json, subprocess, sys
rep, out, rows = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, data = r[0], r[1], r[2]
col = {h: i for i, h in enumerate(hdr)}
def val(name):
    v, u = float(data[col[name]].replace(",", "")), units[col[name]]
    return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u])
j = {"kernel": "ttirt::" + data[col["Kernel Name"]].split("::")[-1].split("(")[0].strip(), "rows_per_launch": rows,
     "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
     "gpu_time_us": float(data[col["gpu__time_duration.sum"]].replace(",", "")),
     "source": "ncu --set full --clock-control none, %s" % rep}
json.dump(j, open(out, "w"), indent=1)
print(j)
