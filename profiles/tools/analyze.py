import csv,re,sys,bisect,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
keys=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__registers_per_thread','lts__t_sector_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_per_inst_issued.ratio']
for i,h in enumerate(hdr):
    if h in keys or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0)>0.3):
        print(f"{h:85s} {units[i]:10s} {vals[i]}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
dev=open('/root/repo/quadrupedal_foothold_planner_b200/csrc/fhp_device.cuh').read().split('\n')
marks=[]
for i,l in enumerate(dev,1):
    if l.startswith('__device__') or l.startswith('template <bool DEFAULT_ONLY>'):
        name=re.findall(r'(\w+)\(',l)
        if name: marks.append((i,name[0]))
cur=None; agg={}; total=0; lines=[]
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>8 and r[0].isdigit():
        try: inst=int(r[7]); smp=int(r[6]); thr=int(r[8])
        except: continue
        ln=int(r[0])
        if cur=='fhp_device.cuh':
            k=bisect.bisect_right([m[0] for m in marks],ln)-1
            key=marks[k][1] if k>=0 else 'top'
        else: key=cur
        agg[key]=agg.get(key,0)+inst; total+=inst
        lines.append((inst,smp,cur,ln,r[1].strip()[:80],thr))
nf=float(sys.argv[2]) if len(sys.argv)>2 else 4194304
print("total inst/foothold",total/nf)
for k,v in sorted(agg.items(),key=lambda x:-x[1])[:16]:
    print(f"{v/total*100:5.1f}%  {v/nf:7.1f} inst/foothold  {k}")
lines.sort(reverse=True)
tots=sum(l[1] for l in lines)
print("-- top lines by instructions")
for a in lines[:14]:
    print(f"{a[0]/total*100:5.1f}% smp {a[1]/tots*100:4.1f}% thr/inst {a[5]/max(a[0],1):5.1f} {a[2]}:{a[3]}  {a[4]}")
lines.sort(key=lambda x:-x[1])
print("-- top lines by stall samples")
for a in lines[:10]:
    print(f"smp {a[1]/tots*100:4.1f}% inst {a[0]/total*100:5.1f}% {a[2]}:{a[3]}  {a[4]}")
if len(sys.argv)>4:
    lo,hi=int(sys.argv[3]),int(sys.argv[4])
    print("-- lines",lo,hi)
    for a in sorted(lines,key=lambda x:x[3]):
        if a[2]=='fhp_device.cuh' and lo<=a[3]<=hi:
            print(f"{a[0]/nf:6.2f} i/f smp {a[1]/tots*100:4.1f}% thr {a[5]/max(a[0],1):5.1f} :{a[3]}  {a[4]}")
if len(sys.argv)>5:
    fn=sys.argv[5]
    print("-- file",fn)
    for a in sorted(lines,key=lambda x:x[3]):
        if a[2]==fn and a[0]/nf>0.05:
            print(f"{a[0]/nf:6.2f} i/f smp {a[1]/tots*100:4.1f}% thr {a[5]/max(a[0],1):5.1f} :{a[3]}  {a[4]}")
