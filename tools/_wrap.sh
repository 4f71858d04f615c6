mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --no-cpu --steps 2 --warmup 1 --epsilon 0.08 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:search_fast -s 2 -c 1 -f -o gpurun_out/s11_prof python bench.py --no-cpu --steps 2 --warmup 1 --epsilon 0.08 > gpurun_out/ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --steps 2 --warmup 1 --epsilon 0.08 > gpurun_out/ncu_l.log 2>&1
tail -c 700 gpurun_out/bench_ours.json
