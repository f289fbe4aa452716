"""Per-stage instruction counters of one bench step, from an ncu CSV launch list:

    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/counters.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline [--config cX]
    python scripts/collect_stage_counters.py gpurun_out/counters.csv c2 512 > profiles/stage_counters.json

bench.py turns these counts (instructions per frame do not depend on the run) and its live stage times into issue-slot
utilisation: issue_frac = warp instructions / (SMs x 4 schedulers x SM clock x stage time)."""
import csv
import json
import sys

STAGE = (("trace", ("k_first_hit", "k_bounce", "k_compact", "k_tree_level")), ("accumulate", ("k_accumulate", "k_reduce_samples")),
         ("post", ("k_post_", "k_psf_", "k_peak_masks", "k_envelope_lerp", "k_transpose", "k_log_compress", "k_image_max")))


def stage_of(kernel: str):
    for name, keys in STAGE:
        if any(k in kernel for k in keys):
            return name
    return None


def main():
    path, config, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    launches = {}
    order = []
    for r in rd:
        i = int(r["ID"])
        if i not in launches:
            launches[i] = {"kernel": r["Kernel Name"], "grid": r.get("Grid Size", "")}
            order.append(i)
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e-3
        elif r["Metric Unit"] in ("ns", "nsecond"):
            v *= 1e-6
        elif r["Metric Unit"] in ("s", "second"):
            v *= 1e3
        launches[i][r["Metric Name"]] = v
    # steps: a step ends with the last post kernel before the next trace kernel
    steps, cur = [], []
    prev_stage = None
    for i in order:
        st = stage_of(launches[i]["kernel"])
        if st is None:
            continue
        if st == "trace" and prev_stage == "post":
            steps.append(cur); cur = []
        cur.append((st, launches[i]))
        prev_stage = st
    if cur:
        steps.append(cur)
    # the first full-size step after the first (cold) one
    def size(step):
        return sum(l.get("smsp__inst_executed.sum", 0.0) for _, l in step)
    big = max(size(s) for s in steps)
    cands = [s for s in steps if size(s) > 0.9 * big]
    step = cands[1] if len(cands) > 1 else cands[0]
    out = {"config": config, "frames_per_launch": frames, "source": path, "n_steps_seen": len(steps), "stages": {}, "kernels": []}
    for name, _ in STAGE:
        ls = [l for st, l in step if st == name]
        out["stages"][name] = {"warp_inst": sum(l.get("smsp__inst_executed.sum", 0.0) for l in ls),
                               "thread_inst": sum(l.get("smsp__thread_inst_executed.sum", 0.0) for l in ls),
                               "ms_under_ncu": sum(l.get("gpu__time_duration.sum", 0.0) for l in ls), "launches": len(ls)}
    for st, l in step:
        out["kernels"].append({"stage": st, "kernel": l["kernel"][:80], "warp_inst": l.get("smsp__inst_executed.sum", 0.0),
                               "thread_inst": l.get("smsp__thread_inst_executed.sum", 0.0), "ms_under_ncu": l.get("gpu__time_duration.sum", 0.0)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
