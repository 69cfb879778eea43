"""Dev tool: per-source-line stall samples / executed instructions of one ncu report (source page, SASS+CUDA view).

usage: python tools/ncu_lines.py report.ncu-rep [file-substring] [top-n]
"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hidx = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and "Instructions Executed" in r]
    grand = 0
    tables = []
    for k, h in enumerate(hidx):
        hdr = rows[h]
        path = rows[h - 2][1] if h >= 2 else "?"
        end = hidx[k + 1] - 2 if k + 1 < len(hidx) else len(rows)
        isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
        samp, ex, src = collections.Counter(), collections.Counter(), {}
        st = collections.defaultdict(collections.Counter)
        for r in rows[h + 1:end]:
            try:
                ln = int(r[0])
            except (ValueError, IndexError):
                continue
            src[ln] = r[1]
            try:
                samp[ln] += int(r[isamp])
            except ValueError:
                pass
            try:
                ex[ln] += int(r[iex])
            except ValueError:
                pass
            for n, c in stall_cols.items():
                try:
                    st[ln][n] += int(r[c])
                except ValueError:
                    pass
        grand += sum(samp.values())
        tables.append((path, samp, ex, src, st))
    print("total samples", grand)
    for path, samp, ex, src, st in tables:
        print(f"== {path}: {sum(samp.values())} samples ({100.0 * sum(samp.values()) / max(grand, 1):.1f} %), "
              f"{sum(ex.values())} warp instructions")
        if want not in path:
            continue
        for ln, n in samp.most_common(top):
            tops = ", ".join(f"{k[6:]}:{v}" for k, v in st[ln].most_common(3) if v)
            print(f"  {ln:5d} {n:7d} {100.0 * n / max(grand, 1):5.1f}% ex {ex[ln]:9d} | {tops} | {src[ln][:110]}")


if __name__ == "__main__":
    main()
