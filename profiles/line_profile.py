import csv,re,sys,subprocess
from collections import defaultdict
def sass_lines(sassfile, kernel_sub):
    """list of (opcode text, file, line) for the kernel whose .text section contains kernel_sub"""
    out=[]; cur=None; inside=False; f=None; ln=None
    for line in open(sassfile):
        m=re.match(r'\s*\.section\s+\.text\.(\S+?),',line)
        if m:
            inside = kernel_sub in m.group(1); continue
        if not inside: continue
        m=re.match(r'\s*//## File "([^"]+)", line (\d+)',line)
        if m: f=m.group(1).split('/')[-1]; ln=int(m.group(2)); continue
        m=re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);',line)
        if m: out.append((m.group(1).strip(),f,ln))
    return out
def ncu_counts(rep, kernel_regex, launch_index=0):
    o=subprocess.run(['ncu','-i',rep,'--page','source','--csv','-k','regex:'+kernel_regex],capture_output=True,text=True).stdout
    rows=list(csv.reader(o.splitlines()))
    # multiple launches concatenated: split at 'Kernel Name' rows
    launches=[]; cur=[]
    for r in rows:
        if r and r[0]=='Kernel Name':
            if cur: launches.append(cur)
            cur=[]; continue
        if r and r[0]=='Address': continue
        cur.append(r)
    if cur: launches.append(cur)
    L=launches[launch_index]
    return [(r[1].strip(), int(r[5]), int(r[2])) for r in L if len(r)>5 and r[5].isdigit()]
if __name__=='__main__':
    rep,kre,ksub=sys.argv[1:4]
    sl=sass_lines('/tmp/cubin/geo.sass',ksub); nc=ncu_counts(rep,kre)
    print('sass',len(sl),'ncu',len(nc))
    agg=defaultdict(lambda:[0,0]); tot=0
    for (txt,f,ln),(t2,n,s) in zip(sl,nc):
        agg[(f,ln)][0]+=n; agg[(f,ln)][1]+=s; tot+=n
    src={}
    for f in ('geometric.cu','vkb_math.cuh','vkb_lattice.cuh'):
        src[f]=open('/root/repo/vkit_b200/csrc/'+f).read().splitlines()
    for (f,ln),(n,s) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[4]) if len(sys.argv)>4 else 30]:
        code=src.get(f,[''])[ln-1].strip()[:80] if f in src and ln and ln<=len(src[f]) else ''
        print(f'{n/tot:6.1%} samp {s:6d} {f}:{ln}  {code}')
