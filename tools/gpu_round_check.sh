set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; echo bench rc $?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r8_bench_ref.json 2> gpurun_out/r8_bench_ref.err; echo ref rc $?
python bench.py --steps 3 --warmup 3 > gpurun_out/r8_bench_steps3.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r8_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r8_ncu_bench.log 2>&1
python tools/prof_ntt.py 14 > gpurun_out/r8_prof_ntt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ntt_forward_kernel -c 1 -s 3 -o gpurun_out/r8_ntt_q62 python tools/prof_ntt.py 14 > gpurun_out/r8_ncu_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:boot_kernel -c 1 -s 5 -o gpurun_out/r8_boot_lean python tools/prof_boot.py 740 742 > gpurun_out/r8_ncu_boot.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ballot_validate -c 1 -s 2 -o gpurun_out/r8_wire_validate python tools/prof_ingest.py 131072 > gpurun_out/r8_ncu_wire.log 2>&1
