"""Per-kernel durations (ms) from an `ncu --csv --metrics gpu__time_duration.sum` launch list: name -> [durations]."""
import csv
import sys
from collections import OrderedDict

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    d = OrderedDict()
    for r in rows[1:]:
        if r[mi] == "gpu__time_duration.sum":
            name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            d.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e6)
    print(path, " ".join("%s=%s" % (k, "/".join("%.2f" % x for x in v)) for k, v in d.items()))
