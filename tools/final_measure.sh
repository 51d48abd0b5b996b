# The round's final evidence, one GPU: gpurun -- 'bash tools/final_measure.sh'   (outputs in gpurun_out/r02zm_*)
# (the reference arm and the teapot capture of r02z stand: neither the oracle nor that kernel changed since)
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02zm_gputests.log 2>&1; tail -2 gpurun_out/r02zm_gputests.log
python bench.py > gpurun_out/r02zm_bench_n1_table.json 2> gpurun_out/r02zm_bench.err
for w in hexagon teapot cow_teddy pumpkin; do python bench.py --workload $w --no-extras --no-cpu-baseline > gpurun_out/r02zm_bench_n1_$w.json 2>> gpurun_out/r02zm_bench.err; done
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02zm_launches_table.csv $B > gpurun_out/ncu_l.log 2>&1
for w in table pumpkin; do $B --workload $w > gpurun_out/plain_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 2 -o gpurun_out/prof_r02zm_$w $B --workload $w > gpurun_out/ncu_$w.log 2>&1; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
