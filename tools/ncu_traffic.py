"""profiles/rNN_traffic.json from an `ncu --set full` report: DRAM bytes of ONE launch of the dominant kernel, which
bench.py reports as roofline.traffic (never a constant in bench.py).  usage: ncu_traffic.py report.ncu-rep chunk_samples out.json"""
import csv, io, json, subprocess, sys

rep, chunk, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, r = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
def val(name):
    v, u = float(r[col[name]].replace(",", "")), units[col[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
kern = r[col["Kernel Name"]]
kern = kern.replace("void ", "").split("(")[0]
d = {"kernel": kern, "chunk_samples": chunk, "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
     "gpu_time_us_under_ncu": float(r[col["gpu__time_duration.sum"]].replace(",", "")) * (1e-3 if units[col["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0),
     "source": "profiles/" + rep.split("/")[-1].replace(".ncu-rep", ".txt") + " (ncu --set full --clock-control none, one launch)"}
json.dump(d, open(out, "w"), indent=1)
print(d)
