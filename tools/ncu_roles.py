"""Where the warps of a warp-specialised kernel spend their samples: reads `ncu -i REP --page source --csv --print-source sass`
(one launch), walks the SASS in address order and charges every instruction's warp-state samples either to the mbarrier wait whose
spin loop it belongs to (TRYWAIT ... BPT.TRAP) or to the work section that ends at the next arrive / commit / wait.

    python tools/ncu_roles.py REPORT.ncu-rep [--launch-skip N] [--names off=name,...]
"""
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    skip = "0"
    names = {}
    for i, a in enumerate(sys.argv):
        if a == "--launch-skip":
            skip = sys.argv[i + 1]
        if a == "--names":
            for kv in sys.argv[i + 1].split(","):
                k, v = kv.split("=")
                names[int(k, 0)] = v
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    heads = [i for i, r in enumerate(rows) if "# Samples" in r]
    h = heads[min(int(skip), len(heads) - 1)]        # one table per launch in the report
    hdr = rows[h]
    i_s, i_src = hdr.index("# Samples"), hdr.index("Source")
    i_addr = hdr.index("Address")
    tab = []
    for r in rows[h + 1:]:
        if len(r) != len(hdr) or r[i_addr] == "Address":
            break
        tab.append(r)
    tab.sort(key=lambda r: int(r[i_addr], 16))
    total = sum(int(r[i_s] or 0) for r in tab)

    def label(text):
        m = re.search(r"\+0x([0-9a-f]+)\]", text)
        off = int(m.group(1), 16) if m else 0
        best = max((k for k in names if k <= off), default=None)
        return f"{names[best]}[{(off - best) // 8}]" if best is not None else hex(off)

    print(f"total samples {total}")
    acc, in_wait, wait_name = 0, False, ""
    for r in tab:
        s, t = int(r[i_s] or 0), r[i_src]
        if "SYNCS.PHASECHK" in t:
            if not in_wait:
                if acc:
                    print(f"  work                      {acc:7d} {100 * acc / total:5.1f} %")
                acc, in_wait, wait_name = 0, True, label(t)
            acc += s
            continue
        if in_wait:
            acc += s
            if "BPT.TRAP" in t:
                print(f"  WAIT {wait_name:20s} {acc:7d} {100 * acc / total:5.1f} %")
                acc, in_wait = 0, False
            continue
        acc += s
        if "SYNCS.ARRIVE" in t or "UTCBAR" in t:
            print(f"  work -> {('arrive ' + label(t)) if 'ARRIVE' in t else 'commit':17s} {acc:7d} {100 * acc / total:5.1f} %")
            acc = 0
        elif "UTCHMMA" in t and acc > 200:
            pass
    if acc:
        print(f"  tail                      {acc:7d} {100 * acc / total:5.1f} %")


if __name__ == "__main__":
    main()
