"""Summarise an .ncu-rep (read here, without a GPU): key throughput metrics and warp-stall shares."""
import csv, json, subprocess, sys
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__grid_size', 'launch__block_size', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__icc_request_hit_rate.pct',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
stall = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
res = []
for r in rows[2:]:
    d = {}
    for k in keep:
        if k in idx:
            d[k] = (r[idx[k]] + ' ' + units[idx[k]]).strip()
    tot = sum(float(r[idx[h]] or 0) for h in stall) or 1
    d['stall_pct'] = {h.split('stalled_')[-1]: round(100 * float(r[idx[h]] or 0) / tot, 1) for h in stall
                      if float(r[idx[h]] or 0) / tot > 0.01}
    res.append(d)
txt = json.dumps(res, indent=1)
if out:
    open(out, 'w').write(txt)
print(txt)
