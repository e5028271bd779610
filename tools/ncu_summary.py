import csv,sys,subprocess,json
WANT=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__block_size','launch__grid_size','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed']
def summarize(rep):
    out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr,units=rows[0],rows[1]
    idx=[hdr.index(w) for w in WANT if w in hdr]
    res=[]
    for r in rows[2:]:
        res.append({hdr[i]+(' ['+units[i]+']' if units[i] else ''): r[i] for i in idx})
    return res
if __name__=='__main__':
    for rep in sys.argv[1:]:
        print('##',rep)
        for d in summarize(rep):
            print(json.dumps(d))
