# one-GPU end-of-round session: parity tests, smoke, the default bench line, the reference arm, the ncu launch list
bash tools/run_gpu_tests.sh tests/test_gpu_*.py
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_final_ref.json 2> gpurun_out/bench_r02_final_ref.err; echo ref rc=$?
python bench.py --profile --steps 2 --warmup 6 > gpurun_out/prof_plain.log 2>&1; echo plain rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 640 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --profile --steps 2 --warmup 6 > gpurun_out/ncu_launch.log 2>&1; echo ncu1 rc=$?
