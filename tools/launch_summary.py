"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel family."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg, cnt, tot = {}, {}, 0.0
    for r in rd:
        if len(r) <= vi:
            continue
        full = r[ki]
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3}.get(r[ui], 1.0)
        name = re.sub(r'\(.*', '', full)
        m = re.search(r'qattention_kernel<(\d+), (\d+), \w+, \w+, (\w+)>', full)
        if m:
            name = f"qd::qattention_kernel<DQ={m.group(1)},DV={m.group(2)},SM16={m.group(3)}>"
        m = re.search(r'gemm_i8_kernel<(-?\d+)>', full)
        if m:
            name = f"qd::gemm_i8_kernel<mode {m.group(1)}>"
        agg[name] = agg.get(name, 0.0) + v
        cnt[name] = cnt.get(name, 0) + 1
        tot += v
    print(f"total {tot / 1e3:.3f} ms over {sum(cnt.values())} launches ({path})")
    fam = collections.defaultdict(float)
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  n={cnt[k]:4d}  {k}")
        fam['gemm' if 'gemm_i8' in k else 'attention' if 'qattention' in k else 'groupnorm' if 'gn_' in k else
            'other elementwise'] += v
    print("families:", {k: f"{v / 1e3:.2f} ms ({100 * v / tot:.1f}%)" for k, v in fam.items()})


if __name__ == "__main__":
    main(sys.argv[1])
