#!/usr/bin/env python
"""ncu `--metrics gpu__time_duration.sum --csv` log -> compact launch list + per-kernel share table (profiles/)."""
import csv, re, sys, collections

def main(src, dst):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]]
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"<.*", "", short).split("::")[-1]
        rows.append((int(r[ix["ID"]]), short, r[ix["Grid Size"]], r[ix["Block Size"]], float(r[ix["Metric Value"]].replace(",", ""))))
    tot = sum(r[4] for r in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        agg[r[1]][0] += 1
        agg[r[1]][1] += r[4]
    with open(dst, "w") as f:
        f.write(f"# source: {src}; {len(rows)} launches, {tot/1e6:.3f} ms summed (ncu: cold-cache, serialised -- compare shares)\n")
        f.write("# kernel,launches,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"# {k},{n},{t/1e3:.1f},{t/tot:.4f}\n")
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows:
            f.write(f'{r[0]},{r[1]},"{r[2]}","{r[3]}",{r[4]:.0f}\n')
    print(f"{len(rows)} launches, {tot/1e6:.3f} ms -> {dst}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
