"""Aggregate an `ncu --page source --csv` export by CUDA source line.

    python profiles/by_line.py <source.csv> <object-or-cubin with -lineinfo> <mangled kernel name substring> [kernel index in csv]

The SASS addresses of the csv are matched (relative to the kernel's first instruction) with `nvdisasm -g` line records of
the same build.  Prints, per source line, the share of executed warp instructions and of stall samples.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_map(obj, name):
    with tempfile.TemporaryDirectory() as d:
        cub = obj
        if not obj.endswith('.cubin'):
            subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
            cub = os.path.join(d, [f for f in os.listdir(d) if f.endswith('.cubin')][0])
        dis = subprocess.run(['nvdisasm', '-g', cub], stdout=subprocess.PIPE, universal_newlines=True).stdout.splitlines()
    out, cur, on = {}, None, False
    for ln in dis:
        if ln.startswith('.text.'):
            on = name in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', ln)
        if m:
            out[int(m.group(1), 16)] = cur
    return out


def main():
    src, obj, name = sys.argv[1:4]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lm = line_map(obj, name)
    part = open(src).read().split('"Kernel Name",')[1 + which]
    lines = part.split('\n')
    print(lines[0][:150])
    rd = csv.reader(io.StringIO('\n'.join(lines[1:])))
    h = next(rd)
    rows = [dict(zip(h, r)) for r in rd if len(r) == len(h)]
    base = int(rows[0]['Address'], 16)
    inst = collections.Counter()
    samp = collections.Counter()
    for r in rows:
        key = lm.get(int(r['Address'], 16) - base)
        inst[key] += int(r['Instructions Executed'])
        samp[key] += int(r['# Samples'])
    ti, ts = sum(inst.values()), sum(samp.values())
    print('executed warp instructions %d, samples %d' % (ti, ts))
    text = {}
    for key in inst:
        if key and key[0] not in text:
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gprmax_b200', 'csrc', key[0])
            text[key[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    for key, v in sorted(inst.items(), key=lambda kv: -(kv[1] / ti + samp[kv[0]] / ts)):
        if v / ti < 0.004 and samp[key] / ts < 0.004:
            continue
        t = ''
        if key and text.get(key[0]) and key[1] - 1 < len(text[key[0]]):
            t = text[key[0]][key[1] - 1].strip()[:90]
        print('%-26s %5.1f%% inst %5.1f%% samples  %s' % ('%s:%d' % key if key else '?', 100 * v / ti, 100 * samp[key] / ts, t))


if __name__ == '__main__':
    main()
