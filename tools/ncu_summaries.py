"""Turn ncu artefacts into the committed text summaries under profiles/."""
import collections, csv, io, re, subprocess, sys

def launch_list(path, out):
    f = open(path).read().split('\n')
    start = [i for i, l in enumerate(f) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(f[start:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')
        v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as o:
        o.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" python bench.py --steps 2 --warmup 3\n')
        o.write(f'# window = the 2 timed steps (NVTX range pushed by bench.py; cold-cache, serialised launches: compare SHARES, not absolutes)\n')
        o.write(f'# {len(rows)} launches, {tot:.1f} us total\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write('%-48s n=%4d  us=%10.1f  share=%.3f\n' % (k[:48], v[0], v[1], v[1] / tot))

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum']

def full(path, out, title):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index('Kernel Name')
    with open(out, 'w') as o:
        o.write(f'# {title}\n# from {path} (ncu --set full --clock-control none --import-source on)\n')
        for r in data:
            o.write(f'\n== {r[ki][:90]}\n')
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    o.write(f'  {w:72s} {r[i]:>16s} {units[i]}\n')
            stalls = [(float(r[i].replace(",", "") or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for i, h in enumerate(hdr)
                      if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and r[i]]
            o.write('  top stall reasons (pc samples): ' + ', '.join(f'{n}={int(v)}' for v, n in sorted(stalls, reverse=True)[:6]) + '\n')

if __name__ == '__main__':
    launch_list('gpurun_out/r1_launches.csv', 'profiles/r1_ncu_launch_list_bench.txt')
    full('gpurun_out/r1_prof_conv_tc.ncu-rep', 'profiles/r1_ncu_conv_tc_bench.txt',
         'k_conv_tc inside bench.py: launches 28..33 of the timed steps = two layer3 bottlenecks (1x1 1024->256 [pair], 3x3 256->256 [pair+slab], 1x1 256->1024 +residual [pair + residual-MMA])')
    full('gpurun_out/r1_prof_post.ncu-rep', 'profiles/r1_ncu_postprocess_bench.txt', 'post-process kernels inside bench.py (res101@550 network output, B=64)')
