#!/bin/bash
# A/B runs of tuning builds on the GPU box: scripts/ab_bench.sh TAG WORKLOAD name1 name2 ...  ("base" = the shipped library)
TAG=$1; WL=$2; shift 2
for n in "$@"; do
  if [ "$n" = base ]; then unset SM_B200_LIB; else export SM_B200_LIB=$PWD/mystereomatching_b200/libsm_b200_$n.so; fi
  python bench.py --workload $WL --no-cpu --no-stream --steps 10 --warmup 3 > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
done
python - "$TAG" "$@" <<'PY'
import json, sys
tag = sys.argv[1]
for n in sys.argv[2:]:
    try:
        d = json.load(open("gpurun_out/%s_%s.json" % (tag, n)))
        print(n, round(d["ms_per_step"], 3), {k: v["ms_per_frame"] for k, v in d["stages"].items()}, d.get("parity", {}).get("pct_identical"), d["quality"])
    except Exception as e:
        print(n, "ERR", e); print(open("gpurun_out/%s_%s.err" % (tag, n)).read()[-1500:])
PY
