#!/bin/bash
# Round-2 closing run on one B200: the whole GPU test suite, smoke(), then both bench arms without a profiler.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1_ours.json 2> gpurun_out/r02f_bench_n1_ours.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1_reference.json 2> gpurun_out/r02f_bench_n1_reference.err
tail -c 300 gpurun_out/r02f_bench_n1_ours.json
