mkdir -p gpurun_out
python bench.py --no-cpu --steps 4 --warmup 1 --epsilon 0.08 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'search_|scan_|prepare_rows' -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --steps 4 --warmup 1 --epsilon 0.08 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-300
