"""Dynamic opcode mix of one captured kernel: `ncu --page source --print-source sass` of a
.ncu-rep (captured with --import-source on) aggregated per SASS opcode, with the share of the
warp-stall samples.  Usage: python scripts/opcode_mix.py REP OUT [warp_steps]  (warp_steps: divide
the counts by it to get instructions per attempted warp-step)."""
import collections
import csv
import io
import re
import subprocess
import sys

FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}


def main(rep, out, warp_steps=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    cnt, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= ix:
            continue
        s = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        op = s.split()[0].split(".")[0]
        cnt[op] += int(r[ix])
        smp[op] += int(r[ismp])
    tot, ts = sum(cnt.values()), sum(smp.values())
    w = float(warp_steps) if warp_steps else None
    with open(out, "a") as fh:
        fh.write(f"\n# dynamic opcode mix ({rows[0][1] if len(rows[0]) > 1 else ''})\n")
        fh.write(f"# warp-level instructions executed: {tot}" + (f" = {tot / w:.1f} per attempted warp-step" if w else "") + "\n")
        f64 = sum(cnt[o] for o in FP64)
        fh.write(f"# fp64-pipe instructions (DFMA+DMUL+DADD+DSETP): {f64}" + (f" = {f64 / w:.1f} per attempted warp-step" if w else "")
                 + f" ({100 * f64 / tot:.1f} % of all)\n")
        fh.write(f"{'opcode':10s} {'executed':>14s} {'share':>7s} {'per step':>9s} {'stall samples':>14s}\n")
        for op, n in cnt.most_common(30):
            fh.write(f"{op:10s} {n:14d} {100 * n / tot:6.2f}% {(n / w if w else 0):9.1f} {100 * smp[op] / ts:13.1f}%\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
