for ch in 8 2; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_|memo_clear" --launch-skip 72 --csv --log-file gpurun_out/r2ag_$ch.csv python profiles/scripts/launches.py 1000000000 3 wp split_chunks=$ch > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open("gpurun_out/r2ag_$ch.csv", errors="replace")))
i0=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; h=rows[i0]; iK=h.index("Kernel Name"); iV=h.index("Metric Value"); iU=h.index("Metric Unit")
t=collections.Counter(); c=collections.Counter()
for r in rows[i0+1:]:
    if len(r)<=iV: continue
    v=float(r[iV].replace(",","")); u=r[iU]; ns=v*{"ns":1,"us":1e3,"ms":1e6}.get(u,1)
    n=r[iK].split("(")[0].replace("void ","")[:40]; t[n]+=ns; c[n]+=1
print("chunks $ch:", {k:(c[k], round(t[k]/1e6,3)) for k in t})
PY
done
for ch in 8 4 2 1; do echo "== chunks $ch"; timeout 200 python profiles/scripts/launches.py 1000000000 2 wp timing=1 split_chunks=$ch 2>&1 | grep timing | tail -1; done
