python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu --epsilon 0.08 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['recall_at_10'], d['config']['overflow_queries_per_step'])"
( python tools/workloads.py hamming --n 1000000; NGTGPU_SO=$PWD/ngt_b200/libngtgpu_w1.so python tools/workloads.py hamming --n 1000000; python tools/workloads.py gist; python tools/workloads.py u8 ) > gpurun_out/workloads2.jsonl 2> gpurun_out/workloads2.err
python - <<'PY'
import json
for l in open('gpurun_out/workloads2.jsonl'):
    d=json.loads(l); s=d['search']
    print(d['workload'], d['n'], 'qps %.0f'%s['qps'], 'eps',s['epsilon'],'rec',s['recall_at_10'],'kernel_ms',s['kernel_ms'],'ms/batch',s['ms_per_batch'],'frac',s['frac_of_measured_hbm_peak'],'ovf',s['overflow_queries'])
PY
tail -2 gpurun_out/workloads2.err
