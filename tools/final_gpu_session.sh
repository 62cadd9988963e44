bash tools/run_gpu_tests.sh tests/test_gpu_*.py
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r02_final_ref.json 2> gpurun_out/bench_r02_final_ref.err; echo ref rc=$?
python bench.py --profile --steps 2 --warmup 6 > gpurun_out/prof_plain.log 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 640 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --profile --steps 2 --warmup 6 > gpurun_out/ncu_launch.log 2>&1; echo ncu1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|head_kernel|mlp3_reduce_kernel|mlp3_tc_bwd2|mlp3_tc_fwd2" -s 1000 -c 5 -f -o gpurun_out/top5_r02 python bench.py --profile --steps 1 --warmup 6 > gpurun_out/ncu_full.log 2>&1; echo ncu2 rc=$?
