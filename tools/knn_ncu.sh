#!/bin/bash
# per-kernel times of one exhaustive kNN pass under ncu (serialised, cold caches: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_launches.csv python tools/knn_probe.py "$@" > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/knn_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e6 if u=="ns" else v/1e3 if u=="us" else v
    k=r[ki][:60]
    if "knn_tc" in k or "scan" in k:
        agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
for k,(c,v) in agg.items(): print("%-62s n=%4d ms=%.3f"%(k,c,v))
PY
