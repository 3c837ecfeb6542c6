import csv,sys,subprocess
from collections import defaultdict, Counter
def launches(path):
    rows=list(csv.reader(open(path)))
    hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
    cols=rows[hdr]; ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
    agg=defaultdict(list)
    for r in rows[hdr+1:]:
        if len(r)>vi: agg[r[ki].split('(')[0][:60]].append(float(r[vi].replace(',','')))
    tot=sum(sum(v) for v in agg.values())
    for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
        print(f'{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us share={sum(v)/tot:6.1%}')
def raw(path, want):
    out=subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr=rows[0]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:50])
        for w in want:
            if w in hdr: print(f'   {w:62s} {r[hdr.index(w)]}')
if __name__=='__main__':
    launches(sys.argv[1])
    raw(sys.argv[2], ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','launch__grid_size'])
