cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 ncu --cache-control none --clock-control none --replay-mode application --kernel-name-base demangled -k regex:'minmax_qparams|QuantizeOp' --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --csv --log-file gpurun_out/observer_pair.csv python tools/profile_observer.py > gpurun_out/observer_pair.log 2>&1; echo rc=$?
tail -3 gpurun_out/observer_pair.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/observer_pair.csv', errors='replace')) if len(r)>10]
h=rows[0]; ki,mi,vi,ui=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('Metric Unit')
idi=h.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[idi], r[ki][:60]), {})[r[mi]]=r[vi]+' '+r[ui]
for k,v in cur.items(): print(k, v)
PY
