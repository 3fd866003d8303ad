"""Summarise an ncu report: headline metrics + instruction/sample share per SASS region."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = list(csv.reader(subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout.splitlines()))
h,u,v = raw[0],raw[1],raw[2]
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:
    if k in h: print(f"{k:75s} {v[h.index(k)]:>16s} {u[h.index(k)]}")
rows = list(csv.reader(subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout.splitlines()))
hdr = rows[1]; data = rows[2:]
iS, iSm, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[iI]) for r in data); ts = sum(int(r[iSm]) for r in data)
stall_cols = [i for i,c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
print("total warp instr", tot, "samples", ts, "sass", len(data))
for b in range(0,len(data),step):
    seg=data[b:b+step]
    e=sum(int(r[iI]) for r in seg); sm=sum(int(r[iSm]) for r in seg)
    if sm/ts>0.02 or e/tot>0.02:
        ops=collections.Counter(); st=collections.Counter()
        for r in seg:
            m=re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS]); ops[m.group(2).split('.')[0]]+=int(r[iI])
            for i in stall_cols:
                try: st[hdr[i][6:]]+=int(r[i])
                except: pass
        print(f"[{b:5d}] instr={100*e/tot:5.1f}% samples={100*sm/ts:5.1f}% ops", dict((k,round(100*c/max(e,1))) for k,c in ops.most_common(5)), "stalls", dict(st.most_common(4)))
