import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hi]; kn = H.index("Kernel Name"); mv = H.index("Metric Value")
for r in rows[hi + 1:]:
    if len(r) > mv: print(r[kn][:90], r[mv])
