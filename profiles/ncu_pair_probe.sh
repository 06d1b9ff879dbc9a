for cfg in "2 8" "1 4"; do set -- $cfg; GPB_PAIR_LAG=$1 GPB_PAIR_XCHUNK=$2 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_update_pair -s 20 -c 3 --csv --log-file gpurun_out/k_pair_$1_$2.csv python profiles/quick_bench.py 300 40 > /dev/null 2>&1; done
GPB_NO_PAIR=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_update_tma -s 40 -c 4 --csv --log-file gpurun_out/k_nopair.csv python profiles/quick_bench.py 300 40 > /dev/null 2>&1
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/k_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]; 
    out={}
    for r in rows[1:]:
        d=dict(zip(hdr,r)); out.setdefault(d['ID'],{})[d['Metric Name']]=(d['Metric Value'],d['Metric Unit'])
    print(f)
    for k,v in out.items(): print('  ',k,v)
PY
