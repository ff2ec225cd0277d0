set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_ba_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_ba_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ba_smoke.log 2>&1; tail -2 gpurun_out/r02_ba_smoke.log
python bench.py > gpurun_out/r02_ba_bench_c3.json 2> gpurun_out/r02_ba_bench_c3.err
for w in c1 c2 c4 c3cen; do python bench.py --workload $w --no-stream > gpurun_out/r02_ba_bench_$w.json 2> gpurun_out/r02_ba_bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 120 --csv --log-file gpurun_out/r02_ba_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream > gpurun_out/r02_ba_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cbca_pass" -s 8 -c 4 -o gpurun_out/r02_ba_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-stream > gpurun_out/r02_ba_ncu2.log 2>&1
ncu -i gpurun_out/r02_ba_prof.ncu-rep --page raw --csv > gpurun_out/r02_ba_prof_raw.csv 2>/dev/null
python - <<'PY'
import json
for w in ("c3","c1","c2","c4","c3cen"):
    try:
        d=json.load(open("gpurun_out/r02_ba_bench_%s.json"%w)); print(w, round(d["ms_per_step"],3), d.get("fps"), d["e2e"].get("fps"), d["roofline"]["frac"], d.get("parity"), d["clocks"])
    except Exception as e: print(w,"ERR",e)
PY
ls -la gpurun_out/r02_ba_prof.ncu-rep
