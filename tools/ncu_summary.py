"""Summarise an ncu launch list (csv) and a full report (--page raw --csv) into small tables for profiles/."""
import collections, csv, re, sys, json

def launches(path):
    rows=[r for r in csv.reader(open(path)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=collections.defaultdict(lambda:[0,0.0])
    for r in rows[1:]:
        try: v=float(r[vi].replace(',',''))
        except ValueError: continue
        name=re.sub(r'\(.*','',r[ki]); name=re.sub(r'.*::','',name).replace(', ','_').replace(',','_')
        agg[name][0]+=1; agg[name][1]+=v
    tot=sum(v[1] for v in agg.values())
    out=['kernel,launches,total_ns,share']
    for k,v in sorted(agg.items(), key=lambda x:-x[1][1]):
        out.append('%s,%d,%.0f,%.4f'%(k,v[0],v[1],v[1]/tot))
    return '\n'.join(out)

WANT=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
      'lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
      'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
      'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
      'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
      'smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum',
      'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
      'sm__cycles_active.avg','sm__cycles_active.max','sm__cycles_elapsed.avg',
      'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']

def full(path):
    rows=list(csv.reader(open(path)))
    hdr=rows[0]; units=rows[1]
    idx=[(w,hdr.index(w)) for w in WANT if w in hdr]
    ki=hdr.index('Kernel Name')
    out=['kernel,'+','.join('%s[%s]'%(w,units[i]) for w,i in idx)]
    for r in rows[2:]:
        name=re.sub(r'\(.*','',r[ki]); name=re.sub(r'.*::','',name).replace(', ','_').replace(',','_')
        out.append(name+','+','.join(r[i].replace(',','') for w,i in idx))
    return '\n'.join(out)

if __name__=='__main__':
    kind,path=sys.argv[1],sys.argv[2]
    print(launches(path) if kind=='launches' else full(path))
