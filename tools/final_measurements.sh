# The measurement pass behind profiles/r02_*: GPU tests, smoke, bench (both arms), ncu launch lists, one --set full capture.
#   gpurun --timeout 1500 -- "bash tools/final_measurements.sh"
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; tail -c 300 gpurun_out/r2_bench21.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench21_ref.json 2>> gpurun_out/r2_bench21.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_bench6.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu6.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_frame25.csv python tools/one_frame.py 3840x2160 3 > /dev/null 2>&1
echo done
