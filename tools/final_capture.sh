# tools/final_capture.sh (GPU box): the measurement set behind profiles/: bench line, per-op and config timings, launch list, one ncu --set full capture of the fused kernel.
# Every ncu pass runs only after the same command has exited 0 without ncu; outputs land in gpurun_out/ (copy the summaries into profiles/ with profiles/analyze.py).
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench.err || exit 1
python tools/bench_ops.py --out gpurun_out/r2f_ops.json > gpurun_out/r2f_ops.log 2>&1
python tools/bench_configs.py --out gpurun_out/r2f_configs.json > gpurun_out/r2f_configs.log 2>&1
python bench.py --filters 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --filters 65536 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pose_fused_kernel -s 3 -c 1 -o gpurun_out/r2f_fused -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -5
