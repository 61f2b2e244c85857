import csv,sys,io,subprocess,re,collections
rep=sys.argv[1]; nbf=int(sys.argv[3]) if len(sys.argv)>3 else 1024*1640
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(io.StringIO(raw))); h=r[0]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_eligible.avg.per_cycle_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
want+=[x for x in h if "issue_stalled" in x and x.endswith("per_issue_active.ratio") and "not_issued" not in x]
for k in want:
    if k in h:
        i=h.index(k); print(f"{k:80s}", [d[i] for d in r[2:]])
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hd=rows[1]; body=rows[2:]
ia=hd.index("Instructions Executed"); isrc=hd.index("Source"); ist=hd.index("Warp Stall Sampling (All Samples)")
ops=collections.Counter()
for b in body:
    m=re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)",b[isrc]); ops[m.group(2)]+=int(b[ia])
print("total warp-instr/bf", sum(ops.values())/nbf)
for k,v in ops.most_common(36): print(f"{k:32s}{v/nbf:8.2f}")
if len(sys.argv)>2:
    open(sys.argv[2],'w').write(src)
