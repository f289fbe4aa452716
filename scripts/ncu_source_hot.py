"""Source-level hot-line table from an ncu report captured with --import-source on (code built with -lineinfo):
per kernel launch, the source lines (file:line) ranked by warp instructions executed, with their share of the
stall samples.  Lines are also summed into named REGIONS (file, first line, last line) given on the command line,
e.g. traversal vs shading.

    python scripts/ncu_source_hot.py gpurun_out/r02a_prof.ncu-rep [--top 25] [--kernel k_bounce]
        [--region name=file.cu:first-last ...]
"""
import argparse
import csv
import io
import subprocess
import sys
from collections import defaultdict


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    launches = []          # (function, {(file, line): [inst, thread_inst, samples, source]})
    cur = None
    cur_file = None
    header = None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            fn = row[1]
            if cur is None or cur[0] != fn or cur[2]:
                pass
            # a new launch starts when the function name changes or the same file shows up again for the same function
            if cur is None or cur[0] != fn or cur_file in cur[3]:
                cur = [fn, defaultdict(lambda: [0, 0, 0, ""]), False, set()]
                launches.append(cur)
            cur[3].add(cur_file)
            continue
        if row[0] == "Line No":
            header = row
            i_inst = header.index("Instructions Executed")
            i_tinst = header.index("Thread Instructions Executed")
            i_samp = header.index("# Samples")
            continue
        if header is None or cur is None:
            continue
        try:
            line = int(row[0])
        except ValueError:
            continue
        if row[2] != "-":          # a SASS row under the source line: already summed in the line's own row
            continue
        try:
            rec = cur[1][(cur_file, line)]
            rec[0] += int(row[i_inst]); rec[1] += int(row[i_tinst]); rec[2] += int(row[i_samp]); rec[3] = row[1].strip()
        except (ValueError, IndexError):
            pass
    return [(l[0], l[1]) for l in launches]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--kernel", default="")
    ap.add_argument("--region", action="append", default=[])
    a = ap.parse_args()
    regions = []
    for r in a.region:
        name, spec = r.split("=")
        f, rng = spec.split(":")
        lo, hi = rng.split("-")
        regions.append((name, f, int(lo), int(hi)))
    for k, (fn, lines) in enumerate(load(a.rep)):
        if a.kernel and a.kernel not in fn:
            continue
        tot_i = sum(v[0] for v in lines.values()) or 1
        tot_t = sum(v[1] for v in lines.values()) or 1
        tot_s = sum(v[2] for v in lines.values()) or 1
        print(f"== launch {k}: {fn[:90]}  warp-instr {tot_i}  thread-instr {tot_t}  samples {tot_s}")
        for name, f, lo, hi in regions:
            ri = sum(v[0] for (ff, ln), v in lines.items() if ff == f and lo <= ln <= hi)
            rs = sum(v[2] for (ff, ln), v in lines.items() if ff == f and lo <= ln <= hi)
            print(f"   region {name:24s} {100.0 * ri / tot_i:5.1f} % instr  {100.0 * rs / tot_s:5.1f} % samples")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[: a.top]:
            print(f"   {100.0 * v[0] / tot_i:5.1f} % instr {100.0 * v[2] / tot_s:5.1f} % samp  {f}:{ln}  {v[3][:110]}")


if __name__ == "__main__":
    sys.exit(main())
