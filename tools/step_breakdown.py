"""print the kernel sequence of the last decode step found in an ncu launch-list csv"""
import csv, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rd=list(csv.DictReader(lines))
rows=[(re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","").replace("mg::",""), float(r["Metric Value"].replace(",",""))/1000.0, r.get("Grid Size","")) for r in rd if r.get("Metric Name")=="gpu__time_duration.sum"]
idx=[i for i,r in enumerate(rows) if r[0].startswith("greedy_select")]
s,e=idx[-2]+1, idx[-1]+1
step=rows[s:e]
print(len(step), "kernels in last step, total us %.1f" % sum(r[1] for r in step))
n=int(sys.argv[2]) if len(sys.argv)>2 else 9
for r in step[:n]: print(f"{r[0]:40s} {r[1]:8.2f} {r[2]}")
print("...")
for r in step[-3:]: print(f"{r[0]:40s} {r[1]:8.2f} {r[2]}")
