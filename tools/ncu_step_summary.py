"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list of
`bench.py --steps 1` into per-kernel shares of ONE forward (the last head1x1 ... tail sequence in the log) and the
average DRAM traffic per launch of the dominant kernel.  Writes JSON (committed under profiles/; bench.py reads
`conv_traffic_bytes_per_launch` from it for roofline.traffic).

    python tools/ncu_step_summary.py gpurun_out/launches_step.csv profiles/r01_ncu_step_summary.json
"""
import csv
import json
import sys


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if r]
    # long format: ID, Process ID, ..., Kernel Name, ..., Metric Name, Metric Unit, Metric Value
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    hdr = rows[hdr_i]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    order = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        lid = int(r[ix["ID"]])
        if lid not in launches:
            launches[lid] = {"name": r[ix["Kernel Name"]]}
            order.append(lid)
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        name = r[ix["Metric Name"]]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
                 "Gbyte": 1e9}.get(unit, 1.0)
        launches[lid][name] = val * scale
    seq = [launches[i] for i in order]
    # last forward: from the last head1x1 launch to the end of the conv run that follows it
    heads = [i for i, l in enumerate(seq) if "head1x1_kernel" in l["name"]]
    start = heads[-1]
    fwd = []
    for l in seq[start:]:
        short = l["name"].split("(")[0].replace("void ", "").replace("sr::", "")
        if fwd and not any(k in short for k in ("conv_tc", "bilinear4_fwd", "head1x1")):
            break
        fwd.append(dict(l, short=short))
    tot = sum(l["gpu__time_duration.sum"] for l in fwd)
    per = {}
    for l in fwd:
        d = per.setdefault(l["short"], dict(launches=0, ms=0.0, dram_read=0.0, dram_write=0.0))
        d["launches"] += 1
        d["ms"] += l["gpu__time_duration.sum"]
        d["dram_read"] += l.get("dram__bytes_read.sum", 0.0)
        d["dram_write"] += l.get("dram__bytes_write.sum", 0.0)
    out = {"source": src, "forward_launches": len(fwd), "forward_ms_under_ncu": round(tot, 3), "kernels": {}}
    conv_bytes = conv_n = 0
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
        out["kernels"][k] = dict(launches=d["launches"], ms=round(d["ms"], 3), share=round(d["ms"] / tot, 4),
                                 dram_bytes_per_launch=round((d["dram_read"] + d["dram_write"]) / d["launches"]))
        if "conv_tc_pair_kernel" in k or "conv_tc_kernel<128" in k or "conv_tc_kernel<(int)128" in k:
            conv_bytes += d["dram_read"] + d["dram_write"]
            conv_n += d["launches"]
    out["conv_traffic_bytes_per_launch"] = round(conv_bytes / max(conv_n, 1))
    out["conv_launches"] = conv_n
    # a whole step of the batched headline workload = everything between two patch-gather launches (one gather per
    # step); when the capture holds one, the per-kernel shares and the conv traffic are taken over the whole step
    gathers = [i for i, l in enumerate(seq) if "patch_gather" in l["name"]]
    if len(gathers) >= 2:
        step = seq[gathers[-2]:gathers[-1]]
        tot_s = sum(l["gpu__time_duration.sum"] for l in step)
        per_s, cb, cn = {}, 0.0, 0
        for l in step:
            short = l["name"].split("(")[0].replace("void ", "").replace("sr::", "")
            d = per_s.setdefault(short, dict(launches=0, ms=0.0, dram=0.0))
            d["launches"] += 1
            d["ms"] += l["gpu__time_duration.sum"]
            d["dram"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
            if "conv_tc_pair_kernel" in short:
                cb += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
                cn += 1
        out["step"] = {"launches": len(step), "ms_under_ncu": round(tot_s, 3),
                       "kernels": {k: dict(launches=d["launches"], ms=round(d["ms"], 3), share=round(d["ms"] / tot_s, 4),
                                           dram_bytes_per_launch=round(d["dram"] / d["launches"]))
                                   for k, d in sorted(per_s.items(), key=lambda kv: -kv[1]["ms"])}}
        out["conv_traffic_bytes_per_launch"] = round(cb / max(cn, 1))
        out["conv_launches"] = cn
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
