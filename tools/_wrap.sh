mkdir -p gpurun_out
( python tools/workloads.py glove; python tools/workloads.py u8 --n 1000000; python tools/workloads.py hamming --n 1000000; python tools/workloads.py gist; python tools/workloads.py u8 ) > gpurun_out/workloads.jsonl 2> gpurun_out/workloads.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --no-cpu --steps 2 --warmup 1 --epsilon 0.08 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --steps 2 --warmup 1 --epsilon 0.08 > gpurun_out/ncu_l.log 2>&1
tail -c 600 gpurun_out/bench_ours.json; wc -l gpurun_out/workloads.jsonl; tail -3 gpurun_out/workloads.err
