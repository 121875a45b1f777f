"""Turn `ncu -i <rep> --page raw --csv --print-units base` output into profiles/dram_traffic.json
(per kernel: launches captured, mean duration, mean DRAM bytes read+written per launch) -- the `roofline.traffic`
source of bench.py.  Usage: python profiles/ncu_traffic.py raw.csv profiles/dram_traffic.json "<capture command>" """
import csv, json, re, sys

ALIAS = {"k_icp_persistent": "icp_fused", "k_gn_persistent": "icp_fused", "k_gen_compact": "gen_surfels",
         "k_compact_update": "compact_scatter"}


def main(src, dst, note):
    rows = list(csv.reader(open(src, newline="")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[hdr]
    col = {n: i for i, n in enumerate(names)}
    out = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kn = re.sub(r"\(.*", "", r[col["Kernel Name"]]).split("::")[-1].replace("void ", "").strip()
        kn = re.sub(r"<.*", "", kn)  # template arguments
        key = ALIAS.get(kn, kn[2:] if kn.startswith("k_") else kn)

        def f(name):
            try:
                return float(r[col[name]].replace(",", ""))
            except (KeyError, ValueError):
                return float("nan")
        e = out.setdefault(key, {"kernel": kn, "rows": []})
        e["rows"].append((f("gpu__time_duration.sum"), f("dram__bytes_read.sum"), f("dram__bytes_write.sum")))
    res = {"source": note, "filter": "launches shorter than 25% of the kernel's longest captured launch are dropped "
                                     "(early-exit launches, e.g. the fallback Gauss-Newton launch that finds done=1)",
           "kernels": {}}
    if "icp_fused" in out:  # the persistent Gauss-Newton kernel runs two jobs per scan: the minimisation (long) and the
        rows = out["icp_fused"]["rows"]  # statistics / recovery job (short) -- bench.py times them as icp_fused / icp_post
        longest = max(r[0] for r in rows)
        post = [r for r in rows if r[0] < 0.5 * longest]
        if post:
            out["icp_post"] = {"kernel": out["icp_fused"]["kernel"] + " (GN_POST job)", "rows": post}
            out["icp_fused"]["rows"] = [r for r in rows if r[0] >= 0.5 * longest]
    for k, e in out.items():
        longest = max(r[0] for r in e["rows"])
        keep = [r for r in e["rows"] if r[0] >= 0.25 * longest]
        e["launches"] = len(keep)
        e["ns"], e["rd"], e["wr"] = (sum(r[i] for r in keep) for i in range(3))
        n = e["launches"]
        res["kernels"][k] = {"kernel": e["kernel"], "launches_captured": n, "avg_us_under_ncu": round(e["ns"] / n / 1e3, 2),
                             "dram_bytes_read_per_launch": round(e["rd"] / n), "dram_bytes_write_per_launch": round(e["wr"] / n),
                             "dram_bytes_per_launch": round((e["rd"] + e["wr"]) / n)}
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res["kernels"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
