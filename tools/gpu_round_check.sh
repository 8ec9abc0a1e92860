# Round-end evidence on one B200 (run through gpurun): parity suite, smoke, both bench arms, launch lists and ncu captures.
set -x
T=${1:-r9}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench rc $?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo ref rc $?
python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_steps3.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_ncu_bench.log 2>&1
python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/${T}_bench_steps3_nosec.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_nosec.csv python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/${T}_ncu_bench_nosec.log 2>&1
python tools/prof_ntt.py 14 > gpurun_out/${T}_prof_ntt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ntt_forward_kernel -c 1 -s 3 -o gpurun_out/${T}_ntt_q62 python tools/prof_ntt.py 14 > gpurun_out/${T}_ncu_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:boot_kernel -c 1 -s 5 -o gpurun_out/${T}_boot_lean python tools/prof_boot.py 740 742 > gpurun_out/${T}_ncu_boot.log 2>&1
