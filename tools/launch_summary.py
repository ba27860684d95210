"""Summarise an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list: time share (and DRAM bytes) per kernel family.
usage: launch_summary.py launches.csv [--traffic out.json]   (--traffic: mean DRAM bytes per GEMM launch for bench.py)"""
import collections
import csv
import json
import re
import sys

UNIT = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6}
BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def family(k):
    return ('gemm' if 'gemm_i8' in k else 'attention' if 'qattention' in k or 'krowsum' in k else
            'groupnorm' if 'gn_' in k else 'other elementwise')


def short(full):
    name = re.sub(r'\(.*', '', full)
    name = re.sub(r'^void ', '', name)
    m = re.search(r'qattention_kernel<(\d+), (\d+), \w+, \w+, (\w+)', full)
    if m:
        name = f"qd::qattention_kernel<DQ={m.group(1)},DV={m.group(2)},SM16={m.group(3)}>"
    m = re.search(r'qattention_smallk_kernel<(\d+), (\d+)', full)
    if m:
        name = f"qd::qattention_smallk_kernel<DQ={m.group(1)},DV={m.group(2)}>"
    m = re.search(r'gemm_i8_kernel<\(?(?:int\))?(-?\d+)>', full)
    if m:
        name = f"qd::gemm_i8_kernel<mode {m.group(1)}>"
    return name


def main(path, traffic_out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    idi, ki, mi, vi, ui = (hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'),
                           hdr.index('Metric Unit'))
    launches = collections.OrderedDict()   # id -> {name, us, bytes}
    for r in rd:
        if len(r) <= vi:
            continue
        L = launches.setdefault(r[idi], {"name": short(r[ki]), "us": 0.0, "bytes": 0.0})
        v = float(r[vi].replace(',', ''))
        if r[mi].startswith('gpu__time_duration'):
            L["us"] += v * UNIT.get(r[ui], 1.0)
        elif r[mi].startswith('dram__bytes'):
            L["bytes"] += v * BYTES.get(r[ui], 1.0)
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for L in launches.values():
        a = agg[L["name"]]
        a[0] += L["us"]; a[1] += 1; a[2] += L["bytes"]
    tot = sum(a[0] for a in agg.values())
    print(f"total {tot / 1e3:.3f} ms over {len(launches)} launches ({path})")
    fam = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for k, (us, n, by) in sorted(agg.items(), key=lambda x: -x[1][0]):
        extra = f"  {by / 1e6:9.1f} MB DRAM ({by / max(us, 1e-9) / 1e6:6.2f} TB/s)" if by else ""
        print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  n={n:4d}  {k}{extra}")
        f_ = fam[family(k)]
        f_[0] += us; f_[1] += by; f_[2] += n
    print("families:", {k: f"{v[0] / 1e3:.2f} ms ({100 * v[0] / tot:.1f}%), {v[1] / 1e9:.2f} GB DRAM, n={v[2]}" for k, v in fam.items()})
    if traffic_out and fam['gemm'][2]:
        g = fam['gemm']
        with open(traffic_out, 'w') as f:
            json.dump({"gemm_dram_bytes_per_launch": g[1] / g[2], "gemm_launches": g[2], "gemm_dram_bytes_per_step": g[1],
                       "gemm_time_share_of_step": g[0] / tot,
                       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the GEMM launches of one timed step ({path})"},
                      f, indent=1)


if __name__ == "__main__":
    out = sys.argv[sys.argv.index('--traffic') + 1] if '--traffic' in sys.argv else None
    main(sys.argv[1], out)
