"""Summarise gpurun_out/adjudication.jsonl (written by tests/helpers.adjudicate_gradients on the GPU box) into a
tracked markdown table: per case and gradient tensor, the error of this library and of the reference CUDA build against
the float64 adjudicator.   python tools/adjudication_summary.py gpurun_out/adjudication.jsonl profiles/r2_adjudication.md"""
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
last = {}
for line in open(src):
    r = json.loads(line)
    if "mean" in next(iter(r["report"].values()))["ours"]:
        last[r["case"]] = r["report"]          # keep the latest run of each case
with open(dst, "w") as f:
    f.write("# Gradient parity adjudicated by float64 (round 2)\n\n"
            "Error of each implementation against `oracle/adjudicator_f64.cu` (the reference's backward algorithm in double with\n"
            "the float32 control flow), per gradient tensor: max / 99.9th percentile / mean absolute error.  `ref` = the\n"
            "unmodified reference CUDA build, largest of three runs (stored run + three live runs for `golden[...]`); `ours` = three\n"
            "runs: largest mean / p99.9, median of the worst element.  Gate (tests/helpers.py):\n"
            "ours <= 1.5x ref (mean), 2x (p99.9), 4x (max; 8x for `golden[...]`), + 1e-5; no element exempt.  Source: the last `pytest -m gpu` run\n"
            "on a B200 (`gpurun_out/adjudication.jsonl`).\n\n")
    order = sorted(last, key=lambda c: (not c.startswith("configs"), not c.startswith("dense"), c))
    for case in order:
        f.write(f"## {case}\n\n| tensor | max abs value | ours max | ours p99.9 | ours mean | ref max | ref p99.9 | ref mean | ours/ref mean |\n|---|---|---|---|---|---|---|---|---|\n")
        for k, v in last[case].items():
            o, r = v["ours"], v["ref"]
            ratio = o["mean"] / r["mean"] if r["mean"] > 0 else float("nan")
            f.write(f"| {k} | {v['scale']:.3g} | {o['max']:.2e} | {o['p999']:.2e} | {o['mean']:.2e} | {r['max']:.2e} | {r['p999']:.2e} | {r['mean']:.2e} | {ratio:.2f} |\n")
        f.write("\n")
print("wrote", dst, len(last), "cases")
