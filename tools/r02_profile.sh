#!/bin/bash
# Round-2 evidence run on one B200: the bench lines of both arms without a profiler, then ncu launch list of one short
# bench run and --set full captures of the dominant kernels. Outputs under gpurun_out/ (summaries go to profiles/).
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_ours.json 2> gpurun_out/r02_bench_n1_ours.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_reference.json 2> gpurun_out/r02_bench_n1_reference.err
# launch list of the headline step (no extras, no cpu leg)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches.csv \
    python bench.py --steps 2 --warmup 3 --extras none --no-cpu > gpurun_out/r02_ncu_step.log 2>&1
# the traversal kernel of the headline (tier-0 launch of the timed steps)
ncu --set full --clock-control none --import-source on -k regex:search_fast_kernel -s 30 -c 1 -o gpurun_out/r02_search_fast_f32 \
    python bench.py --steps 2 --warmup 3 --extras none --no-cpu > gpurun_out/r02_ncu_f32.log 2>&1
# the tensor-core kNN filter (1M x 128 float, k = 128, one batch of three waves)
ncu --set full --clock-control none --import-source on -k regex:knn_tc_filter -s 1 -c 1 -o gpurun_out/r02_knn_tc_sift \
    python tools/knn_probe.py --n 1000000 --k 128 --queries 113664 > gpurun_out/r02_ncu_knn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_tc_filter -s 1 -c 1 -o gpurun_out/r02_knn_tc_glove \
    python tools/knn_probe.py --shape glove --n 1200000 --k 64 --queries 113664 > gpurun_out/r02_ncu_knn2.log 2>&1
# the two-warp lean kernel on uint8 rows
ncu --set full --clock-control none --import-source on -k regex:search_fast_kernel -s 12 -c 1 -o gpurun_out/r02_search_fast_u8 \
    python tools/narrow_probe.py --kind u8 --n 1000000 --settings 0:0:0 > gpurun_out/r02_ncu_u8.log 2>&1
ls -la gpurun_out/*.ncu-rep
