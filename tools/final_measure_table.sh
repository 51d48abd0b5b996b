set -x
python -m pytest tests -m gpu -q > gpurun_out/r02z_gputests.log 2>&1; tail -2 gpurun_out/r02z_gputests.log
python bench.py > gpurun_out/r02z_bench_n1_table.json 2> gpurun_out/r02z_bench.err; tail -c 200 gpurun_out/r02z_bench_n1_table.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_table.csv $B > gpurun_out/ncu_l.log 2>&1
$B > gpurun_out/plain_table.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 2 -o gpurun_out/prof_r02z_table $B > gpurun_out/ncu_table.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
