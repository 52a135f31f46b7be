"""Regenerate profiles/r01_ncu_summary.md from the raw `ncu --page raw --csv` pages kept under profiles/."""
import csv, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
want = {'gpu__time_duration.sum': 'duration',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active': 'tensor pipe cycles active %',
        'dram__bytes_read.sum': 'DRAM read', 'dram__bytes_write.sum': 'DRAM write',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum': 'L2->SM bytes via TMA',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second': 'L2->SM TMA rate',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'L2 throughput %',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'SM throughput %',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue slots active %',
        'launch__registers_per_thread': 'registers/thread', 'launch__shared_mem_per_block_dynamic': 'dyn smem/block',
        'launch__waves_per_multiprocessor': 'waves/SM', 'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps active %'}
caps = (('prof_conv256', 'conv3x3_tc_kernel<256,64,2> (forward / dgrad, 256 px x 256 ch per CTA)'),
        ('prof_conv128', 'conv3x3_tc_kernel<128,64,2>'),
        ('prof_wgrad256', 'conv3x3_wgrad_tc_kernel<256,2> (256 co x 256 ci per CTA)'),
        ('prof_halo32', 'conv3x3_halo_tc_kernel<32,32> (narrow full-resolution layers)'),
        ('prof_loss', 'scribble_loss_fwd / bwd (C = 5 instantiation of the generic kernels; superseded)'),
        ('prof_loss_lean', 'scribble_loss_fwd/bwd_lean_kernel<5, ce> (compile-time C and variant; 12 pairs of 256^2)'),
        ('prof_loss_lean_lvsc', 'scribble_loss_fwd/bwd_lean_kernel<2, ce> (96 pairs of 224^2)'), ('prof_bnbwd', 'bn_bwd_reduce / bn_bwd_apply'))
out = ["# Round-1 ncu captures (`--set full --clock-control none --import-source on`; tests/run_ncu_kernels.sh, tests/run_ncu_final.sh)\n",
       "Raw pages: `profiles/r01_ncu_prof_*_raw.csv`. Durations are cold-cache and serialised (every launch is replayed ~40x).\n"]
for f, title in caps:
    path = os.path.join(ROOT, 'profiles', 'r01_ncu_%s_raw.csv' % f)
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path, errors='ignore')))
    hdr, units = rows[0], rows[1]
    gi, ki = hdr.index('Grid Size'), hdr.index('Kernel Name')
    cols = [(hdr.index(k), v) for k, v in want.items() if k in hdr]
    out.append("\n## %s\n\n| launch | grid | " % title + " | ".join(v for _, v in cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for n, r in enumerate(rows[2:]):
        name = r[ki].split('(')[0].replace('void ', '').replace('pp::', '')
        out.append("| %d %s | %s | " % (n, name[:44], r[gi]) + " | ".join("%s %s" % (r[i][:9], units[i]) for i, _ in cols) + " |\n")
open(os.path.join(ROOT, 'profiles', 'r01_ncu_summary.md'), 'w').write("".join(out))
print("wrote profiles/r01_ncu_summary.md")
