NX=1024 REPS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_small_step|k_small_multi|k_gather_perm|k_scatter_perm|k_zero" --csv --log-file gpurun_out/r2_trsv_launches.csv python tools/prof_trsv.py > gpurun_out/r2_pc_ncu.log 2>&1
tail -2 gpurun_out/r2_pc_ncu.log
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_trsv_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i+1; break
ix={k:j for j,k in enumerate(h)}
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[start:]:
    if len(r)<len(h): continue
    agg[r[ix['Kernel Name']][:70]][0]+=1; agg[r[ix['Kernel Name']][:70]][1]+=float(r[ix['Metric Value']].replace(',',''))
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print('%-72s n=%5d total %9.1f us (2 applications)'%(k,c,t/1e3))
PY
