#!/bin/bash
# Round-end evidence on one box: the whole GPU suite, compute-sanitizer over the kernel tests, the two N=1 bench lines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests.txt
bash tools/sanitize.sh
timeout 900 python bench.py > gpurun_out/r02_bench_fcos_n1.json 2> gpurun_out/r02_bench_fcos_n1.err; tail -c 600 gpurun_out/r02_bench_fcos_n1.json | head -c 300; echo
timeout 900 python bench.py --arch rcnn > gpurun_out/r02_bench_rcnn_n1.json 2> gpurun_out/r02_bench_rcnn_n1.err
python - <<'PY'
import json
for a in ("fcos", "rcnn"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{a}_n1.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(a, round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(r["frac"], 3), "step_flops_frac", round(r["step_flops_frac"], 3),
              "plb", r.get("per_launch_bound"), d["clocks"], {k: (round(v["ms_per_step"], 2) if isinstance(v, dict) and "ms_per_step" in v else None) for k, v in (d.get("extra") or {}).items()})
    except Exception as e:
        print(a, "failed", e)
PY
