"""profiles/rNN_traffic.json from the committed `ncu --set full` captures (profiles/rNN_<kernel>_full.csv, one launch each):
DRAM bytes read + written per launch, which bench.py reports as roofline.traffic.   python tools/make_traffic.py r02"""
import csv, glob, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}
out = {}
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", tag + "_*_full.csv"))):
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "").strip()
        b = sum(float(d[k].replace(",", "")) * UNIT[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out[name] = {"dram_bytes_per_launch": b, "launches_averaged": 1,
                     "gpu_time_s_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * TIME[u["gpu__time_duration.sum"]],
                     "source": "ncu --set full --clock-control none, one launch (profiles/%s)" % os.path.basename(f)}
json.dump(out, open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
