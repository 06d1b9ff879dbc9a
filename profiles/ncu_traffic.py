"""Summarise `ncu --page raw --csv` exports: per captured launch the kernel, duration, DRAM bytes read / written, and a few
counters; writes profiles/traffic.json (what bench.py quotes as roofline.traffic when the kernel sources still hash the same).

    python profiles/ncu_traffic.py <tma300_raw.csv> [<disp300_raw.csv>]"""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_source_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'gprmax_b200', 'csrc')
    # the device code: kernel headers, the item body and the TMA instantiation unit (host-only changes in gpb_core.cu do not
    # alter the kernels a capture was taken from)
    for name in sorted(os.listdir(d)):
        if not (name.endswith('.cuh') or name.endswith('.inc') or name in ('gpb_tma_inst.cu', 'gpb_tma.h')):
            continue
        with open(os.path.join(d, name), 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1] if len(rows) > 1 and rows[1] and rows[1][0] == '' else None
    out = []
    for r in rows[2 if units else 1:]:
        d = dict(zip(hdr, r))
        def f(k):
            try:
                return float(d.get(k, 'nan').replace(',', ''))
            except ValueError:
                return float('nan')
        u = dict(zip(hdr, units)) if units else {}
        def scaled(k):   # ncu prints bytes in K/M/G units depending on magnitude
            v = f(k)
            unit = u.get(k, '')
            return v * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0}.get(unit, 1.0)
        dur = f('gpu__time_duration.sum')
        dur_ns = dur * {'us': 1e3, 'ms': 1e6, 'ns': 1.0, 'usecond': 1e3, 'msecond': 1e6, 'nsecond': 1.0}.get(u.get('gpu__time_duration.sum', 'ns'), 1.0)
        out.append({'kernel': d.get('Kernel Name', '')[:160], 'duration_us': dur_ns / 1e3,
                    'dram_read_bytes': scaled('dram__bytes_read.sum'), 'dram_write_bytes': scaled('dram__bytes_write.sum'),
                    'dram_throughput_pct': f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                    'sm_throughput_pct': f('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
                    'registers_per_thread': f('launch__registers_per_thread'),
                    'l2_hit_rate_pct': f('lts__t_sector_hit_rate.pct'),
                    'issue_active_pct': f('sm__inst_issued.avg.pct_of_peak_sustained_active'),
                    'warps_active_pct': f('sm__warps_active.avg.pct_of_peak_sustained_active')})
    return out


if __name__ == '__main__':
    traffic = {}
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(p):
        traffic = json.load(open(p))
    tma = launches(sys.argv[1])
    for l in tma:
        print(json.dumps(l))
    if tma:
        worst = max(tma, key=lambda l: l['dram_read_bytes'] + l['dram_write_bytes'])
        traffic['bench_300'] = {'kernel_source_hash': kernel_source_hash(), 'file': 'profiles/r2/' + os.path.basename(sys.argv[1]),
                                'dram_bytes_per_launch': worst['dram_read_bytes'] + worst['dram_write_bytes'], 'launches': tma}
    if len(sys.argv) > 2:
        disp = launches(sys.argv[2])
        for l in disp:
            print(json.dumps(l))
        traffic['dispersive_soil_300'] = {'kernel_source_hash': kernel_source_hash(), 'file': 'profiles/r2/' + os.path.basename(sys.argv[2]), 'launches': disp}
    out = os.path.join(ROOT, 'gpurun_out', 'traffic.json') if os.path.isdir(os.path.join(ROOT, 'gpurun_out')) else p
    json.dump(traffic, open(out, 'w'), indent=1)
    print('wrote', out)
