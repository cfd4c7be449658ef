import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); H = r[0]
for n in range(len(r) - 2):
    row = r[2 + n]; print(row[H.index('Kernel Name')][:60])
    v = []
    for i in range(len(H)):
        if 'average_warps_issue_stalled' in H[i]:
            try: v.append((float(row[i].replace(',', '')), H[i][36:]))
            except Exception: pass
    for x in sorted(v, reverse=True)[:6]: print('  %.2f %s' % x)
    for k in H:
        if any(s in k for s in ('bank_conflicts_pipe_lsu_mem_shared.sum', 'wavefronts_mem_shared.sum', 'inst_executed_op_shared', 'sm__cycles_elapsed.avg ', 'lsu_wavefronts.avg.pct', 'sm__inst_executed_pipe_lsu', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_write.sum', 'lts__throughput')):
            print('   ', k, row[H.index(k)])
