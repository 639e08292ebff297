"""Read an `ncu --page source --csv --print-source sass` export: samples per 100-instruction bucket, the hottest instructions
with their dominant stall reasons, every mbarrier wait / named barrier, and the opcode mix of a range.
    python tools/sass_roles.py file.csv [lo hi]"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)


def load(path):
    rows, hdr = [], None
    with open(path) as f:
        for row in csv.reader(f):
            if len(row) > 5 and row[0] == "Address":
                if hdr is not None:
                    break
                hdr = row
                continue
            if hdr and len(row) == len(hdr):
                rows.append(row)
    return hdr, rows


def main():
    hdr, rows = load(sys.argv[1])
    ix = {h: i for i, h in enumerate(hdr)}
    S = [int(r[ix["# Samples"]] or 0) for r in rows]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
    print(len(rows), "instructions,", sum(S), "samples")
    print(" | ".join(f"{i}:{sum(S[i:i + 100])}" for i in range(0, len(rows), 100)))
    for i in sorted(sorted(range(len(rows)), key=lambda i: -S[i])[:25]):
        r = rows[i]
        st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{i:5d} {S[i]:6d} x{r[ix['Instructions Executed']]:>9s} {r[ix['Source']][:84]:84s} {st}")
    for i, r in enumerate(rows):
        src = r[ix["Source"]]
        if "SYNCS.PHASECHK" in src or "BAR." in src:
            print("   sync", i, S[i], src[:80])
    if len(sys.argv) > 3:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
        c, st = collections.Counter(), collections.Counter()
        for r in rows[lo:hi]:
            op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]
            c[op] += int(r[ix["Instructions Executed"]] or 0)
            for h in stalls:
                st[h[6:]] += int(r[ix[h]] or 0)
        t = sum(st.values())
        print("range", lo, hi, "warp-instructions", sum(c.values()), c.most_common(16))
        print("  stalls", [(k, round(100 * v / t, 1)) for k, v in st.most_common(7)])


if __name__ == "__main__":
    main()
