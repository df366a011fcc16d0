"""Summarise an `ncu --page source --csv` dump: samples per warp-role region and top stall sites."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > 10]

    def f(r, k):
        try:
            return float(r[idx[k]])
        except (ValueError, KeyError):
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r, "# Samples") for r in data)
    print("kernel:", rows[0][1][:90])
    print("total samples %d over %d instructions" % (tot, len(data)))
    agg = {}
    for r in data:
        for k in stalls:
            agg[k] = agg.get(k, 0) + f(r, k)
    print("stall mix:", sorted(((k, int(v)) for k, v in agg.items() if v > 0), key=lambda kv: -kv[1])[:8])
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
        s = {k: f(r, k) for k in stalls}
        ss = sorted(((k, int(v)) for k, v in s.items() if v > 0), key=lambda kv: -kv[1])[:2]
        print(r[idx["Address"]][-5:], "%6d" % f(r, "# Samples"), "%9d" % f(r, "Instructions Executed"),
              r[idx["Source"]][:72], ss)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
