#!/bin/bash
# Round-2 closing evidence run on one B200: the bench lines of both arms without a profiler, then the ncu launch list of
# exactly the timed steps and --set full captures of the dominant kernels. Outputs under gpurun_out/ (summaries go to profiles/).
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1_ours.json 2> gpurun_out/r02f_bench_n1_ours.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1_reference.json 2> gpurun_out/r02f_bench_n1_reference.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_step_launches.csv \
    python bench.py --steps 2 --warmup 3 --extras none --no-cpu --profile-region > gpurun_out/r02f_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_fast_kernel -s 30 -c 1 -o gpurun_out/r02f_search_fast_f32 \
    python bench.py --steps 2 --warmup 3 --extras none --no-cpu > gpurun_out/r02f_ncu_f32.log 2>&1
# the tensor-core kNN filter with the query operand streamed (250k x 960 float, k = 64: 46 k-chunks)
ncu --set full --clock-control none --import-source on -k regex:knn_tc_filter -s 1 -c 1 -o gpurun_out/r02f_knn_tc_gist \
    python tools/knn_probe.py --shape gist --n 250000 --k 64 --queries 113664 > gpurun_out/r02f_ncu_knn_gist.log 2>&1
ls -la gpurun_out/*.ncu-rep
