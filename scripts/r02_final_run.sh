# final verification of HEAD on a fresh B200: GPU tests, smoke, the default bench line (what the driver runs), c4
python -m pytest tests -m gpu -x -q > gpurun_out/r02_bx_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_bx_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_bx_smoke.log 2>&1; tail -1 gpurun_out/r02_bx_smoke.log
python bench.py > gpurun_out/r02_bx_bench_c3.json 2> gpurun_out/r02_bx_bench_c3.err; python scripts/bench_brief.py gpurun_out/r02_bx_bench_c3.json | head -3
python bench.py --workload c4 --no-stream > gpurun_out/r02_bx_bench_c4.json 2> gpurun_out/r02_bx_bench_c4.err; python scripts/bench_brief.py gpurun_out/r02_bx_bench_c4.json | head -10
python bench.py --workload c1 --no-stream --no-cpu > gpurun_out/r02_bx_bench_c1.json 2> gpurun_out/r02_bx_bench_c1.err; python scripts/bench_brief.py gpurun_out/r02_bx_bench_c1.json | head -1
for w in c2 c3cen; do python bench.py --workload $w --no-stream --no-cpu > gpurun_out/r02_bx_bench_$w.json 2> gpurun_out/r02_bx_bench_$w.err; python scripts/bench_brief.py gpurun_out/r02_bx_bench_$w.json | head -1; done
