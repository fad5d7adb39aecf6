# The measurement pass behind profiles/r02_*: GPU tests, smoke, bench (both arms), ncu launch lists, one --set full capture.
#   gpurun --timeout 1500 -- "bash tools/final_measurements.sh"
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err; tail -c 300 gpurun_out/r2_bench20.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench20_ref.json 2>> gpurun_out/r2_bench20.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_bench5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu5.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_frame23.csv python tools/one_frame.py 3840x2160 3 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:classify_sweep -s 2 -c 1 -o gpurun_out/r2_cls_v3 python tools/one_frame.py 3840x2160 2 > gpurun_out/r2_ncu_c3.log 2>&1
echo done
