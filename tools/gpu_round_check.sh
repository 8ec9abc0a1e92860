# Round-end evidence on one B200 (run through gpurun): parity suite, smoke, both bench arms, launch lists.
# (the ncu --set full captures of the hot kernels are tools/ncu_round2.sh)
set -x
T=${1:-r9}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc $?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo ref rc $?
python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_steps3.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1
python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/${T}_bench_steps3_nosec.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_nosec.csv python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/${T}_ncu_bench_nosec.log 2>&1
tools/microbench/pipes > gpurun_out/${T}_pipes.txt; tools/microbench/pipes --json >> gpurun_out/${T}_pipes.txt
tools/microbench/bfly > gpurun_out/${T}_bfly.txt
tools/latency --graph > gpurun_out/${T}_latency.txt
bash tools/prof_u32_modes.sh > gpurun_out/${T}_u32_modes.txt 2>&1
