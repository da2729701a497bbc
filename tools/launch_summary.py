"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count / avg / total."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
    d = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0].replace('void ', '').replace('dcb::', '')
        d[(name[:58], r[gi])].append(float(r[vi].replace(',', '')) / 1000)
    tot = sum(sum(v) for v in d.values())
    byk = collections.defaultdict(float)
    for (n, g), v in d.items():
        byk[n.split('<')[0]] += sum(v)
    print(f"total {tot:.0f} us over {sum(len(v) for v in d.values())} launches")
    for n, t in sorted(byk.items(), key=lambda kv: -kv[1])[:20]:
        print(f"  {t:9.0f} us {100 * t / tot:5.1f} %  {n}")
    print()
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:top]:
        print(f"{sum(v):8.0f} us  x{len(v):<3d} avg {sum(v) / len(v):7.1f}  {k[0]} {k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
