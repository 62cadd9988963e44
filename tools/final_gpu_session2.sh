bash tools/run_gpu_tests.sh tests/test_gpu_*.py
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 5 --no-subconfigs > gpurun_out/bench_r02_final_n1c.json 2> gpurun_out/bench_r02_final_n1c.err; echo bench rc=$?
ncu --set full --clock-control none --import-source on -k regex:"gather_kernel|head_kernel|mlp3_reduce_kernel|mlp3_tc_bwd2|mlp3_tc_fwd2" -s 1000 -c 5 -f -o gpurun_out/top5_r02b python bench.py --profile --steps 1 --warmup 6 > gpurun_out/ncu_full.log 2>&1; echo ncu2 rc=$?
