"""ncu launch list with DRAM counters -> profiles/r2_traffic.json (bench.py's roofline.traffic) and a per-kernel text summary.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "timed/" \
        --csv --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 3     (YB_BENCH_QUICK=1)
    python tools/ncu_traffic.py gpurun_out/r2_traffic.csv

Population: EVERY kernel launch of one timed step (NVTX range `timed` pushed by bench.py); traffic = dram read + write bytes per launch."""
import collections, csv, json, re, sys

def main(path, out_json='profiles/r2_traffic.json', out_txt='profiles/r2_ncu_traffic_all_launches.txt'):
    f = open(path).read().split('\n')
    start = [i for i, l in enumerate(f) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(f[start:]))
    per = collections.OrderedDict()
    for r in rows:
        k = int(r['ID'])
        d = per.setdefault(k, {'kernel': re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')})
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        name = r['Metric Name']
        if name == 'gpu__time_duration.sum':
            d['us'] = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
        else:
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            d[name] = v * mult
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d['kernel']]
        a[0] += 1; a[1] += d.get('us', 0.0); a[2] += d.get('dram__bytes_read.sum', 0.0); a[3] += d.get('dram__bytes_write.sum', 0.0)
    tot_us = sum(a[1] for a in agg.values())
    with open(out_txt, 'w') as o:
        o.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --nvtx --nvtx-include "timed/"\n')
        o.write(f'# window = the timed step(s) of bench.py (cold-cache, serialised launches: compare SHARES and BYTES, not absolute times)\n# {len(per)} launches, {tot_us:.1f} us\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write('%-44s n=%4d  us=%9.1f  share=%.3f  dram_read=%9.1f MB  dram_write=%9.1f MB  per-launch=%8.1f MB\n' %
                    (k[:44], a[0], a[1], a[1] / tot_us, a[2] / 1e6, a[3] / 1e6, (a[2] + a[3]) / 1e6 / a[0]))
    conv = [a for k, a in agg.items() if k.startswith('k_conv_tc')]
    n = sum(a[0] for a in conv)
    js = {'kernel': 'k_conv_tc', 'source': f'{out_txt} (ncu dram__bytes_read.sum + dram__bytes_write.sum over ALL {n} k_conv_tc launches of the timed step(s) inside bench.py, res101@550 B=64 fp16)',
          'launches': n, 'dram_bytes_per_launch_avg': sum(a[2] + a[3] for a in conv) / max(1, n),
          'dram_read_bytes_total': sum(a[2] for a in conv), 'dram_write_bytes_total': sum(a[3] for a in conv)}
    fused = [a for k, a in agg.items() if k.startswith('k_bneck_tc')]
    nf = sum(a[0] for a in fused)
    if nf:
        js['fused'] = {'kernel': 'k_bneck_tc', 'launches': nf, 'dram_bytes_per_launch_avg': sum(a[2] + a[3] for a in fused) / nf,
                       'dram_read_bytes_total': sum(a[2] for a in fused), 'dram_write_bytes_total': sum(a[3] for a in fused)}
    json.dump(js, open(out_json, 'w'), indent=1)
    print(json.dumps(js))

if __name__ == '__main__':
    main(*sys.argv[1:])
